// k1_dropon.cu -- K1, the dropon compile: crop/pad the overlay and its alpha mask onto an
// MCU-aligned canvas, convert to the target JPEG's colour space, subsample with the target's
// sampling factors, integer forward DCT, quality-100 quantisation -- one kernel, no libjpeg.
// Replaces mj_compile_dropon (reference: src/dropon.c:325-576), i.e. the two libjpeg encodes
// (src/image.c:257-347) and two coefficient decodes (src/dropon.c:430-576) it performs.
//
// The arithmetic is libjpeg-turbo 3.1.x's integer pipeline restated: jccolor.c
// rgb_ycc_convert, jcsample.c fullsize/h2v1/h2v2/int_downsample, jfdctint.c jpeg_fdct_islow,
// q == 1 quantisation -- bit-exact against the reference build (tests/test_gpu_parity.py::test_k1_compile_bitexact_vs_oracle).
//
// Work unit = one output block of one component, 8 lanes, lane r = sample row r.  The padded
// canvas of src/dropon.c:352-369 is never materialised: canvas pixels are fetched from the
// dropon with crop / block offset applied on the fly and zero outside.
// Output per component: D (overlay coefficients), W (alpha coefficients, DC += 1024,
// src/dropon.c:542) and the per-block class word consumed by K2.
// Roofline: HBM; algorithmic bytes = 6 B/px read (two 3-byte buffers) + 4 B per output coefficient.
#include "mjx_device.cuh"

namespace mjx {

struct K1Comp {
    int16_t  *D;
    int16_t  *W;
    uint32_t *meta;
    int       wb, hb;
    int       he, ve; // horizontal / vertical expansion = max_samp / samp
    int       start;
};

struct K1Params {
    K1Comp         comp[MJX_MAX_COMPONENTS];
    int            ncomp, total_blocks;
    const uint8_t *image3, *alpha3;
    int            dw, dh;
    int            dropon_cs, alpha_cs, target_cs;
    int            boff_x, boff_y, crop_x, crop_y, crop_w, crop_h;
    int            canvas_w, canvas_h;
};

// byte `ch` of canvas pixel (X, Y): the dropon inside the pasted crop, zero elsewhere
__device__ __forceinline__ int canvas_byte(const K1Params &p, const uint8_t *src, int X, int Y, int ch) {
    const int cx = X - p.boff_x, cy = Y - p.boff_y;
    if(cx < 0 || cy < 0 || cx >= p.crop_w || cy >= p.crop_h) return 0;
    return src[((size_t)(cy + p.crop_y) * p.dw + (cx + p.crop_x)) * 3 + ch];
}

// the sample libjpeg's colour converter produces for component c at canvas position (X, Y)
__device__ __forceinline__ int canvas_sample(const K1Params &p, const uint8_t *src, int in_cs, int c, int X, int Y) {
    if(in_cs == MJX_CS_GRAYSCALE) {
        // reference: src/image.c:295-297,331 -- the 3-byte canvas is read as 1 byte per pixel
        // with row stride = width, so sample (X, Y) is canvas byte Y*W + X.
        const long long flat = (long long)Y * p.canvas_w + X;
        const long long pix = flat / 3;
        const int       ch = (int)(flat - pix * 3);
        const int       pY = (int)(pix / p.canvas_w), pX = (int)(pix - (long long)pY * p.canvas_w);
        return canvas_byte(p, src, pX, pY, ch);
    }
    const int cx = X - p.boff_x, cy = Y - p.boff_y;
    int       p0 = 0, p1 = 0, p2 = 0;
    if(cx >= 0 && cy >= 0 && cx < p.crop_w && cy < p.crop_h) {
        const uint8_t *px = src + ((size_t)(cy + p.crop_y) * p.dw + (cx + p.crop_x)) * 3;
        p0 = px[0], p1 = px[1], p2 = px[2];
    }
    if(in_cs == MJX_CS_RGB && p.target_cs != 2 /* JCS_RGB */)
        return c == 0 ? rgb_to_y(p0, p1, p2) : c == 1 ? rgb_to_cb(p0, p1, p2) : rgb_to_cr(p0, p1, p2);
    return c == 0 ? p0 : c == 1 ? p1 : p2;
}

// one row of 8 downsampled, level-shifted samples of block (bx, by)
__device__ __forceinline__ void sample_row(const K1Params &p, const K1Comp &kc, const uint8_t *src, int in_cs, int c,
                                           int bx, int by, int r, int *s) {
    const int he = kc.he, ve = kc.ve, n = he * ve;
#pragma unroll
    for(int x = 0; x < 8; x++) {
        const int X0 = (bx * 8 + x) * he, Y0 = (by * 8 + r) * ve;
        int       sum = 0;
        for(int j = 0; j < ve; j++)
            for(int i = 0; i < he; i++) sum += canvas_sample(p, src, in_cs, c, X0 + i, Y0 + j);
        // jcsample.c: h2v1 bias 0,1,0,1..; h2v2 bias 1,2,1,2..; otherwise round-half-up box mean
        const int bias = (he == 2 && ve == 1) ? (x & 1) : (he == 2 && ve == 2) ? 1 + (x & 1) : n / 2;
        s[x] = (sum + bias) / n - 128;
    }
}

__device__ __forceinline__ void fdct_block(int (&s)[8], int r, unsigned mask) {
    fdct8_islow<0>(s);
    transpose8(s, r, mask);
    fdct8_islow<1>(s);
    transpose8(s, r, mask);
#pragma unroll
    for(int i = 0; i < 8; i++) s[i] = quant_q1(s[i]);
}

static constexpr int kThreads = 256;

__global__ void __launch_bounds__(kThreads) k1_compile_kernel(const K1Params p) {
    const int r = threadIdx.x & 7;
    const int b = blockIdx.x * (kThreads / 8) + (threadIdx.x >> 3);
    if(b >= p.total_blocks) return;
    int c = 0;
#pragma unroll
    for(int i = 1; i < MJX_MAX_COMPONENTS; i++)
        if(i < p.ncomp && b >= p.comp[i].start) c = i;
    const K1Comp  &kc = p.comp[c];
    const int      bi = b - kc.start;
    const int      by = bi / kc.wb, bx = bi - by * kc.wb;
    const unsigned mask = group_mask();

    int s[8];
    sample_row(p, kc, p.image3, p.dropon_cs, c, bx, by, r, s);
    fdct_block(s, r, mask);
    st_row(kc.D + (size_t)bi * 64 + r * 8, row_pack(s));

    sample_row(p, kc, p.alpha3, p.alpha_cs, c, bx, by, r, s);
    fdct_block(s, r, mask);
    if(r == 0) s[0] += 1024; // reference: src/dropon.c:542
    st_row(kc.W + (size_t)bi * 64 + r * 8, row_pack(s));
    const uint32_t meta = classify_alpha(s, r, mask);
    if(r == 0) kc.meta[bi] = meta;
}

// ---------------------------------------------------------------------------------------
// the strip kernel: what every 3-byte-per-pixel dropon takes (RGB / YCC, i.e. everything but the reference's garbled
// grayscale case above).  A CTA owns a strip of kStripMcus MCUs of one MCU row:
//   1. its threads fetch the strip's pixels ONCE (three colour bytes + the alpha byte each, crop / block offset / zero fill
//      applied on the fly), convert each pixel ONCE to the target's components and park the four sample planes in shared
//      memory -- the per-block kernel above re-read and re-converted every pixel for every component it contributes to;
//   2. 8-lane groups take the strip's blocks from a task list: downsample from shared memory, integer FDCT, q = 1
//      quantisation.  The alpha plane of a sampling class is transformed ONCE and stored to every component of the class
//      (reference: every alpha component is the same plane, src/dropon.c:391-414; with equal sampling factors their
//      coefficients are identical).
// Bit-exact with the per-block kernel (tests/test_gpu_parity.py::test_k1_*).
// ---------------------------------------------------------------------------------------
static constexpr int kStripMcus = 8;

struct K1Strip {
    K1Params p;
    int      mcu_w, mcu_h, mcus_x, mcus_y;
    int      hs[MJX_MAX_COMPONENTS], vs[MJX_MAX_COMPONENTS];    // sampling factors (blocks per MCU)
    int      leader[MJX_MAX_COMPONENTS];                         // first component with the same (he, ve): its alpha blocks serve this one
    int      convert;                                            // 1: RGB dropon onto a non-RGB target (jccolor); 0: bytes as they are
};

__global__ void __launch_bounds__(kThreads) k1_strip_kernel(const K1Strip s) {
    extern __shared__ unsigned char k1_smem[];
    const K1Params &p = s.p;
    const int cols = kStripMcus * s.mcu_w, pitch = cols + 4;
    unsigned char *sp = k1_smem; // [4][mcu_h][pitch]: components 0..2 of the image, then alpha
    const int plane_sz = s.mcu_h * pitch;
    const int mcu_x0 = blockIdx.x * kStripMcus, mcu_y = blockIdx.y;
    const int X0 = mcu_x0 * s.mcu_w, Y0 = mcu_y * s.mcu_h;
    // ---- 1. pixels -> sample planes ----
    for(int i = threadIdx.x; i < s.mcu_h * cols; i += kThreads) {
        const int py = i / cols, px = i - py * cols;
        const int cx = X0 + px - p.boff_x, cy = Y0 + py - p.boff_y;
        int c0 = 0, c1 = 0, c2 = 0, a = 0;
        if(cx >= 0 && cy >= 0 && cx < p.crop_w && cy < p.crop_h) {
            const size_t o = ((size_t)(cy + p.crop_y) * p.dw + (cx + p.crop_x)) * 3;
            c0 = __ldg(p.image3 + o), c1 = __ldg(p.image3 + o + 1), c2 = __ldg(p.image3 + o + 2);
            a = __ldg(p.alpha3 + o);
        }
        if(s.convert) {
            const int y = rgb_to_y(c0, c1, c2), cb = rgb_to_cb(c0, c1, c2), cr = rgb_to_cr(c0, c1, c2);
            c0 = y, c1 = cb, c2 = cr;
        }
        unsigned char *q = sp + py * pitch + px;
        q[0] = (unsigned char)c0, q[plane_sz] = (unsigned char)c1, q[2 * plane_sz] = (unsigned char)c2, q[3 * plane_sz] = (unsigned char)a;
    }
    __syncthreads();
    // ---- 2. blocks ----
    // tasks of the strip: for every component its image blocks, then for every LEADER component its alpha blocks
    const int r = threadIdx.x & 7, grp = threadIdx.x >> 3;
    const unsigned mask = group_mask();
    int nblk[MJX_MAX_COMPONENTS], total = 0;
    for(int c = 0; c < p.ncomp; c++) {
        nblk[c] = kStripMcus * s.hs[c] * s.vs[c];
        total += nblk[c] * (s.leader[c] == c ? 2 : 1);
    }
    for(int t = grp; t < total; t += kThreads / 8) {
        // decode the task
        int c = 0, rest = t, alpha = 0;
        for(; c < p.ncomp; c++) {
            if(rest < nblk[c]) break;
            rest -= nblk[c];
        }
        if(c == p.ncomp) {
            alpha = 1;
            for(c = 0; c < p.ncomp; c++) {
                if(s.leader[c] != c) continue;
                if(rest < nblk[c]) break;
                rest -= nblk[c];
            }
        }
        const K1Comp &kc = p.comp[c];
        const int bw = kStripMcus * s.hs[c];              // blocks per strip row of this component
        const int by_l = rest / bw, bx_l = rest - by_l * bw;
        const int bx = mcu_x0 * s.hs[c] + bx_l, by = mcu_y * s.vs[c] + by_l;
        if(bx >= kc.wb || by >= kc.hb) continue; // strip tail (whole 8-lane group leaves)
        const int he = kc.he, ve = kc.ve, n = he * ve;
        const unsigned char *pl = sp + (alpha ? 3 : c) * plane_sz + (by_l * 8 + r) * ve * pitch + bx_l * 8 * he;
        int v[8];
#pragma unroll
        for(int x = 0; x < 8; x++) {
            int sum = 0;
            for(int j = 0; j < ve; j++)
                for(int i = 0; i < he; i++) sum += pl[j * pitch + x * he + i];
            const int bias = (he == 2 && ve == 1) ? (x & 1) : (he == 2 && ve == 2) ? 1 + (x & 1) : n / 2; // jcsample.c
            v[x] = (sum + bias) / n - 128;
        }
        fdct_block(v, r, mask);
        const size_t bi = (size_t)by * kc.wb + bx;
        if(!alpha) st_row(kc.D + bi * 64 + r * 8, row_pack(v));
        else {
            if(r == 0) v[0] += 1024; // reference: src/dropon.c:542
            const Row8     row = row_pack(v);
            const uint32_t meta = classify_alpha(v, r, mask);
            for(int c2 = c; c2 < p.ncomp; c2++) {
                if(s.leader[c2] != c) continue;
                st_row(p.comp[c2].W + bi * 64 + r * 8, row);
                if(r == 0) p.comp[c2].meta[bi] = meta;
            }
        }
    }
}

// class words for a dropon whose D/W planes were uploaded from the host
__global__ void __launch_bounds__(kThreads) classify_kernel(const int16_t *W, uint32_t *meta, int nblocks) {
    const int r = threadIdx.x & 7;
    const int b = blockIdx.x * (kThreads / 8) + (threadIdx.x >> 3);
    if(b >= nblocks) return;
    int w[8];
    row_unpack(ld_row_keep(W + (size_t)b * 64 + r * 8), w);
    const uint32_t m = classify_alpha(w, r, group_mask());
    if(r == 0) meta[b] = m;
}

__global__ void count_classes_kernel(const uint32_t *meta, int nblocks, unsigned long long *counts) {
    __shared__ unsigned int local[4];
    if(threadIdx.x < 4) local[threadIdx.x] = 0;
    __syncthreads();
    for(int i = blockIdx.x * blockDim.x + threadIdx.x; i < nblocks; i += gridDim.x * blockDim.x)
        atomicAdd(&local[meta_cls(meta[i]) & 3u], 1u);
    __syncthreads();
    if(threadIdx.x < 4 && local[threadIdx.x]) atomicAdd(&counts[threadIdx.x], (unsigned long long)local[threadIdx.x]);
}

cudaError_t launch_k1(cudaStream_t s, const uint8_t *image3, const uint8_t *alpha3, int dw, int dh, int dropon_cs,
                      int target_cs, int boff_x, int boff_y, int crop_x, int crop_y, int crop_w, int crop_h,
                      int canvas_w, int canvas_h, int max_h, int max_v, mjx_dropon *d) {
    K1Params p{};
    p.ncomp = d->view.ncomp;
    p.total_blocks = d->view.total_blocks;
    for(int c = 0; c < p.ncomp; c++) {
        p.comp[c].D = d->D[c];
        p.comp[c].W = d->W[c];
        p.comp[c].meta = d->meta[c];
        p.comp[c].wb = d->view.comp[c].wb;
        p.comp[c].hb = d->view.comp[c].hb;
        p.comp[c].he = max_h / d->view.comp[c].hs;
        p.comp[c].ve = max_v / d->view.comp[c].vs;
        p.comp[c].start = d->view.comp[c].start;
    }
    p.image3 = image3;
    p.alpha3 = alpha3;
    p.dw = dw;
    p.dh = dh;
    p.dropon_cs = dropon_cs;
    p.alpha_cs = target_cs == 2 ? MJX_CS_RGB : MJX_CS_YCC; // reference: src/dropon.c:411-414
    p.target_cs = target_cs;
    p.boff_x = boff_x, p.boff_y = boff_y;
    p.crop_x = crop_x, p.crop_y = crop_y, p.crop_w = crop_w, p.crop_h = crop_h;
    p.canvas_w = canvas_w, p.canvas_h = canvas_h;
    if(p.total_blocks <= 0) return cudaSuccess;
    if(dropon_cs != MJX_CS_GRAYSCALE && max_h >= 1 && max_h <= 4 && max_v >= 1 && max_v <= 4) {
        K1Strip st{};
        st.p = p;
        st.mcu_w = 8 * max_h, st.mcu_h = 8 * max_v;
        st.mcus_x = canvas_w / st.mcu_w, st.mcus_y = canvas_h / st.mcu_h;
        st.convert = (dropon_cs == MJX_CS_RGB && target_cs != 2 /* JCS_RGB */) ? 1 : 0;
        bool ok = st.mcus_x > 0 && st.mcus_y > 0 && st.mcus_y <= 65535 && canvas_w % st.mcu_w == 0 && canvas_h % st.mcu_h == 0;
        for(int c = 0; c < p.ncomp; c++) {
            st.hs[c] = d->view.comp[c].hs, st.vs[c] = d->view.comp[c].vs;
            // the strip addresses a component's blocks as MCU index * sampling factor: needs whole expansion ratios and planes of
            // exactly mcus * factor blocks (what mj_compile_dropon's MCU-aligned canvas gives)
            ok = ok && st.hs[c] * p.comp[c].he == max_h && st.vs[c] * p.comp[c].ve == max_v && p.comp[c].wb == st.mcus_x * st.hs[c] &&
                 p.comp[c].hb == st.mcus_y * st.vs[c];
            st.leader[c] = c;
            for(int c2 = 0; c2 < c; c2++)
                if(p.comp[c2].he == p.comp[c].he && p.comp[c2].ve == p.comp[c].ve && p.comp[c2].wb == p.comp[c].wb && p.comp[c2].hb == p.comp[c].hb) {
                    st.leader[c] = st.leader[c2];
                    break;
                }
        }
        if(ok) {
            const int smem = 4 * st.mcu_h * (kStripMcus * st.mcu_w + 4);
            if(smem > 48 * 1024) {
                cudaError_t e = cudaFuncSetAttribute(k1_strip_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
                if(e != cudaSuccess) return e;
            }
            k1_strip_kernel<<<dim3((st.mcus_x + kStripMcus - 1) / kStripMcus, st.mcus_y), kThreads, smem, s>>>(st);
            return cudaGetLastError();
        }
    }
    const int per = kThreads / 8;
    k1_compile_kernel<<<(p.total_blocks + per - 1) / per, kThreads, 0, s>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_classify(cudaStream_t s, mjx_dropon *d) {
    const int per = kThreads / 8;
    for(int c = 0; c < d->view.ncomp; c++) {
        const int nb = d->view.comp[c].wb * d->view.comp[c].hb;
        if(nb <= 0) continue;
        classify_kernel<<<(nb + per - 1) / per, kThreads, 0, s>>>(d->W[c], d->meta[c], nb);
        cudaError_t e = cudaGetLastError();
        if(e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

// counts_dev: [MJX_MAX_COMPONENTS][4] blocks per class, per component
cudaError_t launch_count_classes(cudaStream_t s, const mjx_dropon *d, unsigned long long *counts_dev) {
    for(int c = 0; c < d->view.ncomp; c++) {
        const int nb = d->view.comp[c].wb * d->view.comp[c].hb;
        if(nb <= 0) continue;
        int grid = (nb + 255) / 256;
        if(grid > 1024) grid = 1024;
        count_classes_kernel<<<grid, 256, 0, s>>>(d->meta[c], nb, counts_dev + 4 * c);
        cudaError_t e = cudaGetLastError();
        if(e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

} // namespace mjx
