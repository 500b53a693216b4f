// k1_lists.cu -- second half of the dropon compile: turn the per-block classes written by K1 into
// what K2 consumes.
//   * two ordered work lists (OPAQUE+U blocks, G blocks), so K2 never touches a transparent block
//     and never mixes classes inside a warp;
//   * for the G blocks only, two compact float arrays in list order: Ds = D * IDCT prescale and
//     A = IDCT2(W) / 255, the pixel-domain alpha.  A is what the reference's 64 DCT-domain
//     products amount to (closed form of src/convolve.c); precomputing it once per
//     compiled dropon removes one of the three 2-D transforms from every blended block.
// One-time work per compiled dropon; ordered (deterministic) stream compaction in three tiny kernels.
#include "mjx_device.cuh"

namespace mjx {

static constexpr int kChunk = 1024;

struct ListParams {
    const uint32_t *meta[MJX_MAX_COMPONENTS];
    int             wb[MJX_MAX_COMPONENTS];
    int             start[MJX_MAX_COMPONENTS];
    int             pad[MJX_MAX_COMPONENTS];  // slots inserted before component c's entries in the generic list
    int             spad[MJX_MAX_COMPONENTS]; // ... in the simple list
    int             ncomp, total_blocks;
};

// 0: not listed (transparent), 1: simple (OPAQUE / U), 2: generic
__device__ __forceinline__ int list_kind(const ListParams &p, int b, uint32_t *entry) {
    if(b >= p.total_blocks) return 0;
    int c = 0;
#pragma unroll
    for(int i = 1; i < MJX_MAX_COMPONENTS; i++)
        if(i < p.ncomp && b >= p.start[i]) c = i;
    const int      bi = b - p.start[c];
    const uint32_t cls = meta_cls(__ldg(p.meta[c] + bi));
    const int      row = bi / p.wb[c];
    *entry = entry_pack(c, row, bi - row * p.wb[c]);
    return cls == CLS_T ? 0 : (cls == CLS_G ? 2 : 1);
}

__global__ void __launch_bounds__(kChunk) list_count_kernel(const ListParams p, uint32_t *chunk_counts) {
    uint32_t  e;
    const int kind = list_kind(p, blockIdx.x * kChunk + threadIdx.x, &e);
    const int ns = __syncthreads_count(kind == 1);
    const int ng = __syncthreads_count(kind == 2);
    if(threadIdx.x == 0) {
        chunk_counts[2 * blockIdx.x] = ns;
        chunk_counts[2 * blockIdx.x + 1] = ng;
    }
}

// exclusive scan over the chunks, in place (a few thousand entries at most: serial per list)
__global__ void list_scan_kernel(uint32_t *chunk_counts, int nchunks) {
    if(threadIdx.x < 2) {
        uint32_t run = 0;
        for(int i = 0; i < nchunks; i++) {
            const uint32_t v = chunk_counts[2 * i + threadIdx.x];
            chunk_counts[2 * i + threadIdx.x] = run;
            run += v;
        }
    }
}

__global__ void __launch_bounds__(kChunk) list_fill_kernel(const ListParams p, const uint32_t *chunk_offsets,
                                                           uint32_t *list_simple, uint32_t *list_generic) {
    __shared__ uint32_t warp_s[kChunk / 32], warp_g[kChunk / 32];
    uint32_t            e = 0;
    const int           kind = list_kind(p, blockIdx.x * kChunk + threadIdx.x, &e);
    const unsigned      bs = __ballot_sync(0xffffffffu, kind == 1), bg = __ballot_sync(0xffffffffu, kind == 2);
    const int           lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if(lane == 0) {
        warp_s[warp] = __popc(bs);
        warp_g[warp] = __popc(bg);
    }
    __syncthreads();
    if(warp == 0) { // exclusive scan of the 32 per-warp counts
        uint32_t vs = warp_s[lane], vg = warp_g[lane];
        uint32_t is = vs, ig = vg;
#pragma unroll
        for(int d = 1; d < 32; d <<= 1) {
            const uint32_t ts = __shfl_up_sync(0xffffffffu, is, d), tg = __shfl_up_sync(0xffffffffu, ig, d);
            if(lane >= d) {
                is += ts;
                ig += tg;
            }
        }
        warp_s[lane] = is - vs;
        warp_g[lane] = ig - vg;
    }
    __syncthreads();
    const unsigned below = (1u << lane) - 1u;
    if(kind == 1) list_simple[chunk_offsets[2 * blockIdx.x] + warp_s[warp] + __popc(bs & below) + p.spad[entry_comp(e)]] = e;
    if(kind == 2) list_generic[chunk_offsets[2 * blockIdx.x + 1] + warp_g[warp] + __popc(bg & below) + p.pad[entry_comp(e)]] = e;
}

// compact float arrays of the generic blocks; 8 lanes per block, lane r = row r
__global__ void __launch_bounds__(256) generic_prepare_kernel(const DropView dv, float *gDs, float *gA, float *gAd) {
    const int r = threadIdx.x & 7;
    const int g = blockIdx.x * 32 + (threadIdx.x >> 3);
    if(g >= dv.n_generic) return;
    const uint32_t  e = __ldg(dv.list_generic + g);
    if(e == 0xffffffffu) return; // padding slot (whole 8-lane group leaves)
    const DropComp &dc = dv.comp[entry_comp(e)];
    const size_t    bi = (size_t)entry_row(e) * dc.wb + entry_col(e);
    int             D[8], W[8];
    row_unpack(ld_row_keep(dc.D + bi * 64 + r * 8), D);
    row_unpack(ld_row_keep(dc.W + bi * 64 + r * 8), W);
    const float    isc[8] = MJX_INV_SCALE_INIT;
    const float    pr = c_inv_scale[r];
    const unsigned mask = group_mask();
    float          ds[8], a[8], dp[8];
#pragma unroll
    for(int i = 0; i < 8; i++) {
        const float s = pr * isc[i];
        ds[i] = (float)D[i] * s;
        dp[i] = ds[i];
        a[i] = (float)W[i] * (s * (1.0f / 255.0f));
    }
    idct8(a);               // lane = vertical frequency, elements = pixel column
    transpose8(a, r, mask); // lane = pixel column, elements = vertical frequency
    idct8(a);               // elements = pixel row
    transpose8(a, r, mask); // lane = pixel row: natural [py][px]
    idct8(dp);              // the overlay's pixels (level-shifted), the same way: what the tensor-core G kernel blends against
    transpose8(dp, r, mask);
    idct8(dp);
    transpose8(dp, r, mask);
    float4 *o = reinterpret_cast<float4 *>(gDs + (size_t)g * 64 + r * 8);
    o[0] = make_float4(ds[0], ds[1], ds[2], ds[3]);
    o[1] = make_float4(ds[4], ds[5], ds[6], ds[7]);
    // A is stored "Q-paired" for k2_generic_kernel: float (8*i + k) * 2 + h  =  A[row 2i + h][col k]
    float *ao = gA + (size_t)g * 64 + (size_t)(r >> 1) * 16 + (r & 1);
    float *po = gAd + (size_t)g * 64 + (size_t)(r >> 1) * 16 + (r & 1);
#pragma unroll
    for(int k = 0; k < 8; k++) {
        ao[2 * k] = a[k];
        po[2 * k] = a[k] * dp[k];
    }
}

cudaError_t launch_build_lists(cudaStream_t s, mjx_dropon *d, uint32_t *chunk_counts_dev, int *launches) {
    const int total = d->view.total_blocks;
    if(total <= 0) return cudaSuccess;
    const int  nchunks = (total + kChunk - 1) / kChunk;
    ListParams p{};
    p.ncomp = d->view.ncomp;
    p.total_blocks = total;
    for(int c = 0; c < p.ncomp; c++) {
        p.meta[c] = d->meta[c];
        p.wb[c] = d->view.comp[c].wb > 0 ? d->view.comp[c].wb : 1;
        p.start[c] = d->view.comp[c].start;
        p.pad[c] = d->generic_pad[c];
        p.spad[c] = d->simple_pad[c];
    }
    cudaError_t e;
    list_count_kernel<<<nchunks, kChunk, 0, s>>>(p, chunk_counts_dev);
    if((e = cudaGetLastError()) != cudaSuccess) return e;
    list_scan_kernel<<<1, 32, 0, s>>>(chunk_counts_dev, nchunks);
    if((e = cudaGetLastError()) != cudaSuccess) return e;
    list_fill_kernel<<<nchunks, kChunk, 0, s>>>(p, chunk_counts_dev, const_cast<uint32_t *>(d->view.list_simple),
                                                const_cast<uint32_t *>(d->view.list_generic));
    if((e = cudaGetLastError()) != cudaSuccess) return e;
    if(launches) *launches += 3;
    if(d->view.n_generic > 0) {
        generic_prepare_kernel<<<(d->view.n_generic + 31) / 32, 256, 0, s>>>(d->view, const_cast<float *>(d->view.gDs),
                                                                              const_cast<float *>(d->view.gA), const_cast<float *>(d->view.gAd));
        if((e = cudaGetLastError()) != cudaSuccess) return e;
        if(launches) *launches += 1;
    }
    return cudaSuccess;
}

int list_chunks(int total_blocks) { return (total_blocks + kChunk - 1) / kChunk; }

} // namespace mjx
