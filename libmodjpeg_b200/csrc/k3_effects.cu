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

// component ends with AC == 0: 8 lanes per block, every lane stores its (zero) row; a CTA owns whole block rows
__global__ void __launch_bounds__(kThreads) k3_rewrite_kernel(const K3Params p, int first_is_zero) {
    const mjx_image_desc_t &im = p.items[blockIdx.y];
    const int c = p.comp, r = threadIdx.x & 7;
    const int wreal = im.wreal[c], hreal = im.hreal[c], stride = im.stride_blocks[c];
    const int q0 = im.q[c][0];
    int16_t  *plane = reinterpret_cast<int16_t *>(im.plane[c]);
    for(int l = blockIdx.x; l < hreal; l += gridDim.x) {
        int16_t *rowp = plane + (size_t)l * stride * 64;
        // four blocks per thread and trip: the DC loads of all four are in flight before the first store
        for(int k0 = threadIdx.x >> 3; k0 < wreal; k0 += 4 * (kThreads / 8)) {
            int dc[4] = {0, 0, 0, 0};
            if(r == 0 && !first_is_zero) {
#pragma unroll
                for(int u = 0; u < 4; u++) {
                    const int k = k0 + u * (kThreads / 8);
                    if(k < wreal) dc[u] = p.dc_compact ? (int)p.dc_compact[(size_t)l * wreal + k] : (int)rowp[(size_t)k * 64];
                }
            }
#pragma unroll
            for(int u = 0; u < 4; u++) {
                const int k = k0 + u * (kThreads / 8);
                if(k >= wreal) break;
                Row8 row;
                row.w[0] = row.w[1] = row.w[2] = row.w[3] = 0;
                if(r == 0) row.w[0] = (uint32_t)fold_dc(p, dc[u], q0) & 0xffffu;
                st_row_stream(rowp + (size_t)k * 64 + r * 8, row);
            }
        }
    }
}

// DC-only pipeline: one thread per block.  A CTA owns block rows l = blockIdx.x, blockIdx.x + gridDim.x, ..;
// its threads walk the columns, four rows at a time so that four independent 2-byte loads are in flight per
// thread (the pass moves one 32-byte DRAM sector each way per 128-byte block: it lives on memory parallelism).
__global__ void __launch_bounds__(kThreads) k3_dc_kernel(const K3Params p) {
    const mjx_image_desc_t &im = p.items[blockIdx.y];
    const int c = p.comp;
    const int wreal = im.wreal[c], hreal = im.hreal[c], stride = im.stride_blocks[c];
    const int q0 = im.q[c][0];
    int16_t  *plane = reinterpret_cast<int16_t *>(im.plane[c]);
    for(int l0 = blockIdx.x * 4; l0 < hreal; l0 += gridDim.x * 4) {
        for(int k = threadIdx.x; k < wreal; k += kThreads) {
            int16_t *bp[4];
            int      dc[4];
#pragma unroll
            for(int u = 0; u < 4; u++) {
                bp[u] = plane + ((size_t)(l0 + u) * stride + k) * 64;
                dc[u] = (l0 + u < hreal) ? (int)*bp[u] : 0;
            }
#pragma unroll
            for(int u = 0; u < 4; u++)
                if(l0 + u < hreal) *bp[u] = (int16_t)fold_dc(p, dc[u], q0);
        }
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
        // CTAs per image: enough for 148 SMs x 8 resident CTAs a few times over; each CTA strides over block rows
        int gx = (148 * 32 + n - 1) / n;
        if(gx < 1) gx = 1;
        if(gx > 1024) gx = 1024;
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
