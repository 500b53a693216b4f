// k3_effects.cu -- K3, the coefficient-only effects, as one fused pass per component.
// Replaces mj_effect_grayscale / _pixelate / _tint / _luminance (reference: src/effect.c:28-222).
//
// An effect pipeline is a short list of steps applied in order to every REAL block of a
// component (width_in_blocks x height_in_blocks; MCU padding blocks are left alone, like the
// reference's loops, src/effect.c:45-50):
//   ZERO      all 64 coefficients := 0                                  (grayscale, per chroma plane)
//   PIXELATE  coefficients 1..63 := 0                                   (pixelate)
//   ADD_DC    DC = (int16)(DC*q0); DC = (int16)(DC+v); clamp +-2047; DC = (int16)(DC/q0)  (tint, luminance)
// Because nothing ever un-zeroes an AC coefficient, a component whose list contains ZERO or
// PIXELATE ends with all AC == 0 and only the DC needs to be read ("rewrite" kernel: 8 lanes per
// block, 2 B read + 128 B written); a list of ADD_DC steps only touches the DC ("dc" kernel: one
// thread per block, 2 B read + 2 B written; the DRAM sector floor is 32 B each way).
// All integer, bit-exact including int16 wrap-around.  Roofline: HBM.
#include "mjx_device.cuh"

namespace mjx {

static constexpr int kMaxOps = 16;

struct K3Params {
    const mjx_image_desc_t *items;
    int                     comp;
    int                     nops;   // steps that apply to `comp`, in order
    int                     op[kMaxOps];
    int                     value[kMaxOps];
    const int16_t          *dc_compact; // rewrite kernel, single staged image: DCs come from this [hreal][wreal] array
};

// reference: src/effect.c:143-153 / 207-217
__device__ __forceinline__ int add_dc(int dc, int q0, int value) {
    int t = wrap16(dc * q0);
    t = wrap16(t + value);
    t = t > 2047 ? 2047 : (t < -2047 ? -2047 : t);
    return wrap16(t / q0);
}

__device__ __forceinline__ int fold_dc(const K3Params &p, int dc, int q0) {
    for(int i = 0; i < p.nops; i++) {
        if(p.op[i] == MJX_FX_ZERO) dc = 0;
        else if(p.op[i] == MJX_FX_ADD_DC) dc = add_dc(dc, q0, p.value[i]);
    }
    return dc;
}

static constexpr int kThreads = 256;
static constexpr int kDcUnroll = 8; // independent DC loads in flight per thread

// Component ends with AC == 0 ("rewrite": pixelate, grayscale).  A warp owns runs of 32 consecutive blocks of one block row
// (4 KB); the two halves of the job are decoupled so that neither waits for the other:
//   gather:  lane i fetches the DC of block i (one 2-byte load = one DRAM sector per block), kDcUnroll runs deep, so a
//            warp has 32 x kDcUnroll sectors in flight before it stores anything;
//   stream:  8 lanes per block write its 128 bytes (the folded DC in the first word, zeros elsewhere) with 128-bit streaming
//            stores, four blocks per instruction; the lane that writes a block's first chunk gets the DC by shuffle.
// Work item = (block row, group of kDcUnroll runs) of image blockIdx.y, dealt to the warps of the image's CTAs.
__global__ void __launch_bounds__(kThreads) k3_rewrite_kernel(const K3Params p, int first_is_zero) {
    const int lane = threadIdx.x & 31;
    const int c = p.comp;
    const mjx_image_desc_t &im = p.items[blockIdx.y];
    const int wreal = im.wreal[c], hreal = im.hreal[c];
    if(im.plane[c] == 0 || wreal <= 0) return;
    const int q0 = im.q[c][0];
    const int groups = (wreal + 32 * kDcUnroll - 1) / (32 * kDcUnroll); // per block row
    const int items = hreal * groups, warps = gridDim.x * (kThreads / 32);
    for(int it = blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5); it < items; it += warps) {
        const int l = it / groups, g = it - l * groups;
        int16_t  *rowp = reinterpret_cast<int16_t *>(im.plane[c]) + (size_t)l * im.stride_blocks[c] * 64;
        const int k_base = g * kDcUnroll * 32;
        int dc[kDcUnroll];
#pragma unroll
        for(int u = 0; u < kDcUnroll; u++) {
            const int k = k_base + u * 32 + lane;
            dc[u] = 0;
            if(!first_is_zero && k < wreal) dc[u] = p.dc_compact ? (int)p.dc_compact[(size_t)l * wreal + k] : (int)__ldcs(rowp + (size_t)k * 64);
        }
#pragma unroll
        for(int u = 0; u < kDcUnroll; u++) {
            const int k0 = k_base + u * 32;
            if(k0 >= wreal) break; // warp-uniform
            const int folded = fold_dc(p, dc[u], q0) & 0xffff;
            int16_t  *runp = rowp + (size_t)k0 * 64 + lane * 8; // lane -> 16-byte chunk (lane & 7) of block (lane >> 3) + 4 i
#pragma unroll
            for(int i = 0; i < 8; i++) {
                const int blk = 4 * i + (lane >> 3);
                const int v = __shfl_sync(0xffffffffu, folded, blk);
                if(k0 + blk < wreal) __stcs(reinterpret_cast<uint4 *>(runp + (size_t)i * 256), make_uint4((lane & 7) == 0 ? (uint32_t)v : 0u, 0u, 0u, 0u));
            }
        }
    }
}

// DC-only pipeline (tint, luminance): one thread per block, kDcUnroll consecutive block ROWS per thread so that its loads are
// independent and all in flight before the first store.  The pass moves one 32-byte DRAM sector each way per 128-byte block: it
// lives on memory parallelism.  Work item = (group of kDcUnroll block rows, 256 columns) of image blockIdx.y.
__global__ void __launch_bounds__(kThreads) k3_dc_kernel(const K3Params p) {
    const int c = p.comp;
    const mjx_image_desc_t &im = p.items[blockIdx.y];
    const int wreal = im.wreal[c], hreal = im.hreal[c], stride = im.stride_blocks[c];
    if(im.plane[c] == 0 || wreal <= 0) return;
    const int q0 = im.q[c][0];
    const int rgroups = (hreal + kDcUnroll - 1) / kDcUnroll, cgroups = (wreal + kThreads - 1) / kThreads;
    for(int it = blockIdx.x; it < rgroups * cgroups; it += gridDim.x) {
        const int rg = it / cgroups, cg = it - rg * cgroups;
        const int k = cg * kThreads + threadIdx.x, l0 = rg * kDcUnroll;
        if(k >= wreal) continue;
        int16_t  *bp = reinterpret_cast<int16_t *>(im.plane[c]) + ((size_t)l0 * stride + k) * 64;
        int       dc[kDcUnroll];
#pragma unroll
        for(int u = 0; u < kDcUnroll; u++) dc[u] = (l0 + u < hreal) ? (int)bp[(size_t)u * stride * 64] : 0;
#pragma unroll
        for(int u = 0; u < kDcUnroll; u++)
            if(l0 + u < hreal) bp[(size_t)u * stride * 64] = (int16_t)fold_dc(p, dc[u], q0);
    }
}

// DC-only pipeline on a COMPACT array of DC values (one int16 per block): what the host-pointer entry point
// stages for one image whose planes live in pageable host memory -- 2 bytes per block cross PCIe instead of 128
__global__ void __launch_bounds__(kThreads) k3_dc_compact_kernel(const K3Params p, int16_t *dc, int nblk, int q0) {
    for(int i = blockIdx.x * kThreads + threadIdx.x; i < nblk; i += gridDim.x * kThreads) dc[i] = (int16_t)fold_dc(p, (int)dc[i], q0);
}

cudaError_t launch_k3_dc_compact(cudaStream_t s, int16_t *dc_dev, int nblk, int q0, int comp, const mjx_effect_op_t *ops, int nops) {
    K3Params p{};
    p.comp = comp;
    for(int i = 0; i < nops; i++) {
        if(ops[i].comp != comp) continue;
        if(p.nops == kMaxOps) return cudaErrorInvalidValue;
        p.op[p.nops] = ops[i].op;
        p.value[p.nops] = ops[i].value;
        p.nops++;
    }
    if(p.nops == 0 || nblk <= 0) return cudaSuccess;
    int grid = (nblk + kThreads - 1) / kThreads;
    if(grid > 148 * 8) grid = 148 * 8;
    k3_dc_compact_kernel<<<grid, kThreads, 0, s>>>(p, dc_dev, nblk, q0);
    return cudaGetLastError();
}

cudaError_t launch_k3(cudaStream_t s, const mjx_image_desc_t *items_dev, int n, int ncomp, const mjx_effect_op_t *ops,
                      int nops, int *launches, const int16_t *const *dc_compact) {
    if(n <= 0) return cudaSuccess;
    for(int c = 0; c < ncomp; c++) {
        K3Params p{};
        p.items = items_dev;
        p.comp = c;
        p.dc_compact = (dc_compact && n == 1) ? dc_compact[c] : nullptr;
        bool rewrite = false;
        for(int i = 0; i < nops; i++) {
            if(ops[i].comp != c) continue;
            if(p.nops == kMaxOps) return cudaErrorInvalidValue;
            p.op[p.nops] = ops[i].op;
            p.value[p.nops] = ops[i].value;
            p.nops++;
            if(ops[i].op == MJX_FX_ZERO || ops[i].op == MJX_FX_PIXELATE) rewrite = true;
        }
        if(p.nops == 0) continue;
        // CTAs per image: enough to fill 148 SMs x 8 resident CTAs a few times over; each CTA strides over the image's work items
        int gx = (148 * 32 + n - 1) / n;
        if(gx < 1) gx = 1;
        if(gx > 512) gx = 512;
        for(int first = 0; first < n; first += 65535) {
            const int cnt = n - first < 65535 ? n - first : 65535;
            p.items = items_dev + first;
            if(rewrite) k3_rewrite_kernel<<<dim3(gx, cnt), kThreads, 0, s>>>(p, p.op[0] == MJX_FX_ZERO ? 1 : 0);
            else k3_dc_kernel<<<dim3(gx, cnt), kThreads, 0, s>>>(p);
            cudaError_t e = cudaGetLastError();
            if(e != cudaSuccess) return e;
            if(launches) (*launches)++;
        }
    }
    return cudaSuccess;
}

} // namespace mjx
