// mjx_internal.cuh -- host-side objects behind the opaque handles of include/mjx.h and the
// parameter blocks handed to the kernels.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <mutex>
#include <string>
#include <vector>

#include "mjx.h"

namespace mjx {

// ---- device-side view of one component of a compiled dropon -----------------------------
struct DropComp {
    const int16_t  *D;    // [hb][wb][64] overlay coefficients (libjpeg q=1 integer DCT)
    const int16_t  *W;    // [hb][wb][64] alpha coefficients, DC += 1024
    const uint32_t *meta; // [hb][wb]     class | (alpha DC << 8)
    int wb, hb;           // blocks
    int hs, vs;           // sampling factors of this component in the target image
    int start;            // index of this component's first block in the flattened block list
};

// Work lists written by K1 so that K2 never visits a transparent block and never mixes classes
// in a warp.  An entry locates one dropon block: comp << 30 | row << 15 | col (dropon-relative),
// row-major within a component, components in order.
struct DropView {
    DropComp        comp[MJX_MAX_COMPONENTS];
    int             ncomp;
    int             total_blocks;
    const uint32_t *list_simple;  // OPAQUE and U blocks; padded per component like list_generic
    const uint32_t *list_generic; // G blocks; every component starts on a multiple of 32, gaps hold 0xffffffff
    int             n_simple, n_generic; // slots (multiples of 32), not blocks
    int             gtile_start[MJX_MAX_COMPONENTS]; // first tile (32 slots) of each component in list_generic
    int             stile_start[MJX_MAX_COMPONENTS]; // ... in list_simple
    const float    *gDs; // [n_generic][64] overlay coefficients * IDCT prescale (natural order)
    const float    *gA;  // [n_generic][64] pixel-domain alpha / 255 = IDCT2(W) / 255, stored Q-paired: (8i + k)*2 + h = A[2i + h][k]
    const float    *gAd; // [n_generic][64] A o IDCT2(D): alpha times the overlay's pixels (minus 128), Q-paired like gA
};

static inline __host__ __device__ uint32_t entry_pack(int comp, int row, int col) {
    return ((uint32_t)comp << 30) | ((uint32_t)row << 15) | (uint32_t)col;
}
static inline __host__ __device__ int entry_comp(uint32_t e) { return (int)(e >> 30); }
static inline __host__ __device__ int entry_row(uint32_t e) { return (int)((e >> 15) & 0x7fffu); }
static inline __host__ __device__ int entry_col(uint32_t e) { return (int)(e & 0x7fffu); }

} // namespace mjx

namespace mjx {
// what a ctx remembers about its device between K2 launches (function attributes are per device)
struct K2Dev {
    bool g_attr = false;
    bool op_attr[2] = {false, false};
    int  g_ctas_per_sm = 0;
};
// batches of at least this many images take the tensor-core G kernel (k2_generic_op.cu): below it the operator pieces
// (16-24 KB per dropon block and launch) outweigh what the fp32 kernel costs
static constexpr int kOpMinImages = 256;
// ... and by default only batches from this size on: the kernel deals batches of 64 images round-robin to three groups of warps,
// so a small batch leaves groups idle (256 images: 2 + 1 + 1 batches, a third of the tensor-core kernel's time is waiting;
// c2 with 256 images: 0.233 ms against 0.200 ms on the fp32 kernel).  From 17 batches on at most 1 in 10 group slots is empty.
static constexpr int kOpDefaultMinImages = 1025;
struct OpView;
} // namespace mjx

// ---- opaque handles -------------------------------------------------------------------
struct mjx_dropon {
    int            device = 0;
    mjx_layout_t   layout{};
    void          *slab = nullptr;  // D, W, meta of every component
    size_t         slab_bytes = 0;
    void          *slab2 = nullptr; // work lists + compact generic-class arrays
    size_t         slab2_bytes = 0;
    mjx::DropView  view{};
    int16_t       *D[MJX_MAX_COMPONENTS] = {};
    int16_t       *W[MJX_MAX_COMPONENTS] = {};
    uint32_t      *meta[MJX_MAX_COMPONENTS] = {};
    long long      counts[4] = {}; // blocks per class, all components
    int            generic_pad[MJX_MAX_COMPONENTS] = {}; // padding slots before each component's part of the generic list
    int            simple_pad[MJX_MAX_COMPONENTS] = {};  // ... of the simple list
    // operator cache of the tensor-core G kernel (k2_generic_op.cu): allocated by the first ctx that runs a large batch
    // with this dropon and used by that ctx only (the cache is rebuilt in stream order when the batch's quantisation
    // tables change, so it cannot serve two streams at once); every other ctx takes the fp32 kernel
    std::atomic<mjx_ctx *> op_owner{nullptr};
    void                  *op_slab = nullptr;
    size_t                 op_bytes = 0;
    mjx::OpView           *op = nullptr;
};

struct mjx_ctx {
    int          device = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr; // own_stream or a borrowed one
    std::string  last_error;
    long long    launches = 0;
    int          sm_count = 0;
    int          strict = 0; // 1: one kernel for every class with the reference's int16 wrap-around
    int          class_mask = 3; // fast path: bit 0 = OPAQUE/U kernel, bit 1 = G kernel (profiling aid, default both)
    int          zero_copy = 1; // batch-host calls on page-locked planes run K2 directly on host memory

    // staging pools for the host-pointer entry points (grown on demand, reused across calls)
    void  *pin = nullptr;
    size_t pin_bytes = 0;
    void  *pin2 = nullptr; // caller-visible page-locked scratch (mjx_ctx_pinned_scratch)
    size_t pin2_bytes = 0;
    void  *dev = nullptr;
    size_t dev_bytes = 0;
    void  *desc_dev = nullptr; // device array of mjx_image_desc_t for staged launches
    size_t desc_bytes = 0;
    void  *scratch = nullptr; // per-launch K2 scratch: work counters
    size_t scratch_bytes = 0;
    void  *dev2 = nullptr;    // caller-visible device scratch (mjx_ctx_device_scratch)
    size_t dev2_bytes = 0;
    void  *huff = nullptr;    // K4 scratch: block bit lengths, the unstuffed stream, piece counts (k4_huffman.cu)
    size_t huff_bytes = 0;

    // extra streams for the pipelined batch-host path
    static const int kPipe = 3;
    cudaStream_t pipe[kPipe] = {};

    // K2 runs its two kernels side by side: the OPAQUE/U kernel on this low-priority stream, forked from and joined
    // back into `stream` with the two events (k2_compose.cu: launch_k2)
    cudaStream_t side_stream = nullptr;
    cudaEvent_t  side_fork = nullptr, side_join = nullptr;
    int          overlap = 1;
    int          k2_tc = 1; // G class of batches of >= kOpMinImages images on the tensor-core kernel: 1 with range check, 2 without, 0 off
    int          k2_op_min_images = mjx::kOpDefaultMinImages; // mjx_ctx_set_tensor_core_min_images
    int          k2_op_pieces = 2;     // fp16 pieces per operator entry (MJX_K2_OP_PIECES: 2 or 3)
    size_t       k2_op_max_bytes = (size_t)4 << 30; // largest operator cache a dropon may get (MJX_K2_OP_MAX_MB)
    mjx::K2Dev   k2dev;
};

namespace mjx {

int  fail(mjx_ctx *ctx, cudaError_t e, const char *what);
int  ensure_pin(mjx_ctx *ctx, size_t bytes);
int  ensure_dev(mjx_ctx *ctx, size_t bytes);
int  ensure_desc(mjx_ctx *ctx, size_t bytes);
int  ensure_scratch(mjx_ctx *ctx, size_t bytes);

// kernel launchers (each returns a cudaError_t from the launch; *launches += kernels launched)
cudaError_t launch_k1(cudaStream_t s, const uint8_t *image3, const uint8_t *alpha3, int dw, int dh, int dropon_cs,
                      int target_cs, int boff_x, int boff_y, int crop_x, int crop_y, int crop_w, int crop_h,
                      int canvas_w, int canvas_h, int max_h, int max_v, mjx_dropon *d);
cudaError_t launch_classify(cudaStream_t s, mjx_dropon *d);
cudaError_t launch_count_classes(cudaStream_t s, const mjx_dropon *d, unsigned long long *counts_dev);
// after classification: fill the work lists and the compact generic-class arrays (slab2 allocated)
cudaError_t launch_build_lists(cudaStream_t s, mjx_dropon *d, uint32_t *chunk_counts_dev, int *launches);
// side (optional): stream + fork/join events for running the OPAQUE/U kernel beside the G kernel
struct K2Side {
    cudaStream_t stream;
    cudaEvent_t  fork, join;
};
struct K2Launch {
    cudaStream_t  stream = nullptr;
    void         *scratch = nullptr; // k2_scratch_bytes(n, view) bytes
    int           strict = 0;
    int           sm_count = 0;
    int           class_mask = 3;
    int           tc = 0;            // G class of large batches: 0 fp32 kernel; 1 tensor-core kernel, coefficient range checked
                                     // (out-of-range blocks go to the fp32 kernel); 2 tensor-core kernel, range vouched for
    const OpView *op = nullptr;      // the dropon's operator cache (nullptr: not available, fp32 kernel)
    const K2Side *side = nullptr;
    K2Dev        *dev = nullptr;
    int          *launches = nullptr;
};
// bytes of scratch one K2 launch over n images needs (work counters + the redo bitmap of the tensor-core kernel)
size_t      k2_scratch_bytes(int n, const DropView &view);
cudaError_t launch_k2(const K2Launch &L, const mjx_image_desc_t *items_dev, int n, const DropView &view, int block_x, int block_y);
// dc_compact (optional, n == 1): per component a device array [hreal][wreal] the rewrite kernel takes the DCs from
cudaError_t launch_k3(cudaStream_t s, const mjx_image_desc_t *items_dev, int n, int ncomp, const mjx_effect_op_t *ops,
                      int nops, int *launches, const int16_t *const *dc_compact = nullptr);
cudaError_t launch_selftest_reciprocal(cudaStream_t s, unsigned long long *mismatches_dev);
// DC-only step list of component `comp` on a compact device array of nblk DC values
cudaError_t launch_k3_dc_compact(cudaStream_t s, int16_t *dc_dev, int nblk, int q0, int comp, const mjx_effect_op_t *ops, int nops);

} // namespace mjx

#define MJX_CUDA(ctx, call)                                        \
    do {                                                           \
        cudaError_t e__ = (call);                                  \
        if(e__ != cudaSuccess) return mjx::fail(ctx, e__, #call);  \
    } while(0)
