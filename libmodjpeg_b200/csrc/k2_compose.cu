// k2_compose.cu -- K2, the masked blend: dequantise, blend with the compiled dropon, requantise
// to the image's own tables.  Replaces mj_compose_with_mask + mj_convolve
// (reference: src/compose.c:237-342, src/convolve.c:29-1099).
//
// Work unit = one 8x8 block of one component of one image, owned by 8 lanes (lane r = row r,
// one 128-bit load/store per plane).  The block's class, written by K1, decides what is touched:
//   T       nothing: the image block is neither loaded nor stored
//   OPAQUE  load D, store trunc(D / q)                      (bit-exact)
//   U       load I and D, one fp32 multiply per coefficient (bit-exact, see mjx_math.cuh)
//   G       load I, D, W; pixel-domain blend: alpha = IDCT(W)/255, Y = DCT(alpha * IDCT(D - I*q))
//           -- the closed form of the reference's 64 sparse DCT-domain products,
//           within +-1 quantisation step of it.
// Roofline: HBM.  Algorithmic bytes per block: T 0, OPAQUE 128, U/G 256 of image traffic, plus
// the compiled dropon once per launch (L2-resident across the images of a batch).
#include "mjx_device.cuh"

namespace mjx {

struct K2Params {
    DropView                drop;
    const mjx_image_desc_t *items;
    int                     block_x, block_y; // dropon origin on the image, in MCUs
};

static constexpr int kThreads = 256;
static constexpr int kBlocksPerCta = kThreads / 8;

__global__ void __launch_bounds__(kThreads) k2_compose_kernel(const K2Params p) {
    const int r = threadIdx.x & 7;
    const int b = blockIdx.x * kBlocksPerCta + (threadIdx.x >> 3);
    if(b >= p.drop.total_blocks) return;

    int c = 0;
#pragma unroll
    for(int i = 1; i < MJX_MAX_COMPONENTS; i++)
        if(i < p.drop.ncomp && b >= p.drop.comp[i].start) c = i;
    const DropComp &dc = p.drop.comp[c];
    const int       bi = b - dc.start;

    const uint32_t meta = __ldg(dc.meta + bi);
    const uint32_t cls = meta_cls(meta);
    if(cls == CLS_T) return;

    const mjx_image_desc_t &im = p.items[blockIdx.y];
    const int l = bi / dc.wb, k = bi - l * dc.wb;
    const int row = p.block_y * dc.vs + l, col = p.block_x * dc.hs + k;
    if(row >= im.rows[c] || col >= im.stride_blocks[c]) return;

    int16_t       *ip = reinterpret_cast<int16_t *>(im.plane[c]) + ((size_t)row * im.stride_blocks[c] + col) * 64 + r * 8;
    const int16_t *dp = dc.D + (size_t)bi * 64 + r * 8;

    int q[8], D[8], out[8];
    float rq[8];
    {
        Row8 qr = ld_row_keep(&im.q[c][r * 8]);
#pragma unroll
        for(int i = 0; i < 8; i++) {
            q[i] = (int)((qr.w[i >> 1] >> ((i & 1) * 16)) & 0xffffu);
            rq[i] = quant_rcp(q[i]);
        }
    }
    row_unpack(ld_row_keep(dp), D);

    if(cls == CLS_OPAQUE) {
#pragma unroll
        for(int i = 0; i < 8; i++) out[i] = tdiv(D[i], rq[i]);
        st_row_stream(ip, row_pack(out));
        return;
    }

    int I[8];
    row_unpack(ld_row_stream(ip), I);

    if(cls == CLS_U) {
        const float w4 = uniform_w4(meta_wdc(meta));
#pragma unroll
        for(int i = 0; i < 8; i++) out[i] = blend_uniform(I[i], D[i], q[i], rq[i], w4);
        st_row_stream(ip, row_pack(out));
        return;
    }

    // ---- generic block -------------------------------------------------------------------
    const unsigned mask = group_mask();
    int W[8], deq[8];
    row_unpack(ld_row_keep(dc.W + (size_t)bi * 64 + r * 8), W);

    float x[8], a[8];
    const float pr = c_inv_scale[r];
    {
        const float isc[8] = MJX_INV_SCALE_INIT;
#pragma unroll
        for(int i = 0; i < 8; i++) {
            deq[i] = wrap16(I[i] * q[i]);
            const float s = pr * isc[i];
            x[i] = (float)(D[i] - deq[i]) * s;
            a[i] = (float)W[i] * (s * (1.0f / 255.0f));
        }
    }
    idct8(x);
    idct8(a);
    transpose8(x, r, mask);
    transpose8(a, r, mask);
    idct8(x);
    idct8(a);
#pragma unroll
    for(int i = 0; i < 8; i++) x[i] *= a[i];
    fdct8(x);
    transpose8(x, r, mask);
    fdct8(x);
    {
        const float fr = c_fwd_scale[r];
        const float fsc[8] = MJX_FWD_SCALE_INIT;
#pragma unroll
        for(int i = 0; i < 8; i++) {
            const float Y = x[i] * (fr * fsc[i]);
            out[i] = tdiv(wrap16(deq[i] + f2i_trunc(Y)), rq[i]);
        }
    }
    st_row_stream(ip, row_pack(out));
}

cudaError_t launch_k2(cudaStream_t s, const mjx_image_desc_t *items_dev, int n, const DropView &view, int block_x,
                      int block_y) {
    if(n <= 0 || view.total_blocks <= 0) return cudaSuccess;
    K2Params p;
    p.drop = view;
    p.block_x = block_x;
    p.block_y = block_y;
    const unsigned gx = (unsigned)((view.total_blocks + kBlocksPerCta - 1) / kBlocksPerCta);
    for(int first = 0; first < n; first += 65535) {
        const int cnt = n - first < 65535 ? n - first : 65535;
        p.items = items_dev + first;
        k2_compose_kernel<<<dim3(gx, (unsigned)cnt), kThreads, 0, s>>>(p);
        cudaError_t e = cudaGetLastError();
        if(e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

} // namespace mjx
