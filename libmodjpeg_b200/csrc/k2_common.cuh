// k2_common.cuh -- definitions shared by the K2 kernels (k2_compose.cu: OPAQUE/U kernel, fp32 G kernel, strict kernel;
// k2_generic_tc.cu: the tensor-core G kernel).
#pragma once

#include "mjx_device.cuh"

namespace mjx {

struct FastParams {
    DropView                drop;
    const mjx_image_desc_t *items;
    unsigned int           *counter; // work-stealing counter of the generic kernel
    int                     n;       // images
    int                     block_x, block_y;
    int                     images_per_item;
    // blocks the tensor-core G kernel left to the fp32 kernel: word tile * n + image, bit = slot within the tile
    unsigned int           *redo_mask;
    unsigned int           *redo_count;
};

// ---- the tensor-core G kernel of large batches (k2_generic_op.cu) ---------------------------------------------------
static constexpr int kOpGroups = 3; // groups of four compute warps per CTA (three shared-memory stages and two accumulators each)

// device pointers into the operator cache of a compiled dropon
struct OpView {
    unsigned char *B;       // [n_generic][np][8192] operator pieces, fp16, stored as SWIZZLE_128B shared-memory images
    float         *K;       // [n_generic][64] L(D): the overlay's share of the blend term
    float         *q512;    // [MJX_MAX_COMPONENTS][64] 512 q (the staged fp16 operand is I / 512)
    float         *rq;      // [MJX_MAX_COMPONENTS][64] biased reciprocals of the tables
    uint16_t      *key;     // [MJX_MAX_COMPONENTS][64] the tables the cache was built for
    int           *info;    // [MJX_MAX_COMPONENTS] log2 of the operator's scale, -1: component not served (q > 255)
    int           *rebuild; // [MJX_MAX_COMPONENTS] set by the prepare kernel when the tables changed
    int            np;      // fp16 pieces per operator entry (2 or 3)
};

struct OpParams {
    DropView                drop;
    OpView                  op;
    const mjx_image_desc_t *items;
    uint4                  *table; // [ncomp][n] plane address, stride, rows of the images that take this path
    unsigned int           *redo_mask;
    unsigned int           *redo_count;
    int                     n;
    int                     block_x, block_y;
};

// bytes of the operator cache of a dropon with n_generic list slots; with v != nullptr: v->B holds the slab's base on
// entry and every pointer of *v is set on return
size_t      op_cache_bytes(int n_generic, int np, OpView *v);
// prepare (address table, table comparison) + operator build (no-op when the tables are unchanged) + the kernel;
// attr_set[2] caches the per-device function attributes
cudaError_t launch_k2_generic_op(cudaStream_t s, const OpParams &p, int sm_count, bool check, bool *attr_set, int *launches);

static constexpr int kGStages = 2;
// Shared-memory blocks are PADDED by one 16-byte chunk (stride 144 B for int16 blocks, 272 B for
// float blocks): lane t reading chunk c of "its" block t hits bank group (t + c) mod 8, so the
// thread-per-block 128-bit accesses are conflict-free AND every address is lane base + immediate
// (an XOR swizzle costs a LOP3 + IADD per access: ~100 issue slots per block).
static constexpr int kInStride = 144, kF32Stride = 272;
static constexpr int kInBytes = 32 * kInStride; // one image's 32 blocks
static constexpr int kQRawBytes = 128;          // its quantisation table as stored (64 x uint16)
static constexpr int kAddrBytes = 32 * 8;       // global addresses of the 32 blocks
static constexpr int kStageBytes = kInBytes + kQRawBytes + kAddrBytes;
static constexpr int kTabBytes = 3 * 64 * 4; // q * prescale, q, biased 1/q as floats (current image)
static constexpr int kWarpBytes = kGStages * kStageBytes + kTabBytes;
static constexpr int kTileHalf = 32 * kF32Stride; // A (Q-paired) or Ds of the tile
static constexpr int kTileBytes = 2 * kTileHalf;
static constexpr int g_smem(int warps) { return kTileBytes + warps * kWarpBytes; }

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
// 16 bytes global -> shared; nbytes == 0 zero-fills without touching `src` (no branch for absent blocks)
__device__ __forceinline__ void cp_async16(unsigned dst, const void *src, unsigned nbytes = 16u) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(nbytes) : "memory");
}
// the same past L1 (L2 only): a streaming gather must not depend on L1 lines for its misses in flight -- with 227 KB of the SM's
// 256 KB given to shared memory the L1 that is left holds ~200 lines, which caps the bytes in flight per SM (measured:
// k2_generic_op_kernel's gather alone 1.88 ms with .ca)
__device__ __forceinline__ void cp_async16_cg(unsigned dst, const void *src, unsigned nbytes = 16u) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(nbytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// sign-extend with PRMT, convert with the full-rate I2FP.F32.S32 (the compiler's I2F.S16 issues at 1/4 rate)
// (PTX prmt replicates the sign of a byte when bit 3 of its selector nibble is set; the
// __byte_perm() intrinsic masks that bit off, hence the inline asm)
__device__ __forceinline__ float s16lo(uint32_t w) {
    int v;
    asm("prmt.b32 %0, %1, 0, 0x9910;" : "=r"(v) : "r"(w));
    return (float)v;
}
__device__ __forceinline__ float s16hi(uint32_t w) { return (float)((int32_t)w >> 16); }
__device__ __forceinline__ F2 s16pair(uint32_t w) { return f2(s16lo(w), s16hi(w)); }

// forward AAN scale of the pair (row r; cols 2j, 2j+1), indexed 4r + j
struct FwdScale2 {
    float2 v[32];
};
static __constant__ FwdScale2 c_fwd2 = {{
#define MJX_F(r, a, b) {(float)(r * a), (float)(r * b)}
#define MJX_FROW(r)                                                                                              \
    MJX_F(r, 0.35355339059327376, 0.25489778955207959), MJX_F(r, 0.27059805007309851, 0.30067244346752264),     \
        MJX_F(r, 0.35355339059327376, 0.44998811156820786), MJX_F(r, 0.65328148243818826, 1.28145772387075308)
    MJX_FROW(0.35355339059327376), MJX_FROW(0.25489778955207959), MJX_FROW(0.27059805007309851), MJX_FROW(0.30067244346752264),
    MJX_FROW(0.35355339059327376), MJX_FROW(0.44998811156820786), MJX_FROW(0.65328148243818826), MJX_FROW(1.28145772387075308)
#undef MJX_FROW
#undef MJX_F
}};


} // namespace mjx
