// mjx_device.cuh -- device-only helpers shared by the kernels: 128-bit row loads/stores with
// cache hints, the 8-lane register transpose, and block classification.
#pragma once

#include "mjx_internal.cuh"
#include "mjx_math.cuh"

namespace mjx {

// scale factors of the AAN transforms, indexable by a run-time lane id
static __constant__ float c_inv_scale[8] = MJX_INV_SCALE_INIT;
static __constant__ float c_fwd_scale[8] = MJX_FWD_SCALE_INIT;

// the 8 lanes that own one block
__device__ __forceinline__ unsigned group_mask() { return 0xffu << (threadIdx.x & 24); }

// image planes are streamed (read once, written once): evict-first so the compiled dropon and
// the quantisation tables stay in L2 across the images of a batch
__device__ __forceinline__ Row8 ld_row_stream(const int16_t *p) {
    uint4 v = __ldcs(reinterpret_cast<const uint4 *>(p));
    Row8  r;
    r.w[0] = v.x, r.w[1] = v.y, r.w[2] = v.z, r.w[3] = v.w;
    return r;
}
__device__ __forceinline__ void st_row_stream(int16_t *p, const Row8 &r) {
    __stcs(reinterpret_cast<uint4 *>(p), make_uint4(r.w[0], r.w[1], r.w[2], r.w[3]));
}
// compiled dropon / quant tables: read-only path, default (keep) policy
__device__ __forceinline__ Row8 ld_row_keep(const void *p) {
    uint4 v = __ldg(reinterpret_cast<const uint4 *>(p));
    Row8  r;
    r.w[0] = v.x, r.w[1] = v.y, r.w[2] = v.z, r.w[3] = v.w;
    return r;
}
__device__ __forceinline__ void st_row(int16_t *p, const Row8 &r) {
    *reinterpret_cast<uint4 *>(p) = make_uint4(r.w[0], r.w[1], r.w[2], r.w[3]);
}

// 8x8 transpose across the 8 lanes of a block: lane r element i  <->  lane i element r.
// Three butterfly stages, 4 shuffles each.  `r` is the lane's index within its group.
template <typename T>
__device__ __forceinline__ void transpose8(T (&v)[8], int r, unsigned mask) {
#pragma unroll
    for(int d = 4; d >= 1; d >>= 1) {
        const bool up = (r & d) != 0;
#pragma unroll
        for(int i = 0; i < 8; i++) {
            if(i & d) continue;
            T send = up ? v[i] : v[i + d];
            T recv = __shfl_xor_sync(mask, send, d);
            if(up) v[i] = recv;
            else v[i + d] = recv;
        }
    }
}

// class of an alpha block from its rows (lane r holds W row r, DC already += 1024)
__device__ __forceinline__ uint32_t classify_alpha(const int *w, int r, unsigned mask) {
    int ac = 0;
#pragma unroll
    for(int i = 0; i < 8; i++) ac |= (r == 0 && i == 0) ? 0 : w[i];
    const bool any_ac = __any_sync(mask, ac != 0);
    const int  dc = __shfl_sync(mask, w[0], 0, 8);
    uint32_t   cls = any_ac ? CLS_G : (dc == 0 ? CLS_T : (dc == 2040 ? CLS_OPAQUE : CLS_U));
    return meta_pack(cls, dc);
}

} // namespace mjx
