// emul.cpp -- TEST AID: compiles the kernels' per-lane arithmetic (libmodjpeg_b200/csrc/mjx_math.cuh)
// with g++ and replays K2's per-block logic with the 8 lanes of a block as a loop, so the
// arithmetic can be checked against the oracle on the CPU-only build box before GPU time is
// spent.  Not part of the product; the product has no CPU path.
#include <stdint.h>
#include <string.h>

#include "mjx_math.cuh"

using namespace mjx;

extern "C" {

// exhaustive check of tdiv() against C's a / q; returns the number of mismatches
long long emul_tdiv_check(int qmin, int qmax) {
    long long bad = 0;
    for(int q = qmin; q <= qmax; q++) {
        const float rq = quant_rcp(q);
        for(int a = -32768; a <= 32767; a++)
            if(tdiv(a, rq) != a / q) bad++;
    }
    return bad;
}

static int g_generic_v2 = 0;
void emul_set_generic_v2(int on) { g_generic_v2 = on; }
void emul_generic_block_v2(int16_t *I, const int16_t *D, const int16_t *W, const uint16_t *q);

static void transpose(float m[8][8]) {
    for(int i = 0; i < 8; i++)
        for(int j = i + 1; j < 8; j++) {
            float t = m[i][j];
            m[i][j] = m[j][i];
            m[j][i] = t;
        }
}

// class word exactly as classify_alpha() computes it
uint32_t emul_classify(const int16_t *W) {
    bool any_ac = false;
    for(int i = 1; i < 64; i++) any_ac |= W[i] != 0;
    const int dc = W[0];
    return meta_pack(any_ac ? CLS_G : (dc == 0 ? CLS_T : (dc == 2040 ? CLS_OPAQUE : CLS_U)), dc);
}

// K2 on one block, in place on I; returns the class
int emul_compose_block(int16_t *I, const int16_t *D, const int16_t *W, const uint16_t *q) {
    const uint32_t meta = emul_classify(W);
    const uint32_t cls = meta_cls(meta);
    if(cls == CLS_T) return (int)cls;
    float rq[64];
    for(int i = 0; i < 64; i++) rq[i] = quant_rcp(q[i]);
    if(cls == CLS_OPAQUE) {
        for(int i = 0; i < 64; i++) I[i] = (int16_t)tdiv(D[i], rq[i]);
        return (int)cls;
    }
    if(cls == CLS_U) {
        const float w4 = uniform_w4(meta_wdc(meta));
        for(int i = 0; i < 64; i++) I[i] = (int16_t)blend_uniform(I[i], D[i], q[i], rq[i], w4);
        return (int)cls;
    }
    if(g_generic_v2) {
        emul_generic_block_v2(I, D, W, q);
        return (int)cls;
    }
    float x[8][8], a[8][8];
    int   deq[64];
    for(int r = 0; r < 8; r++)
        for(int i = 0; i < 8; i++) {
            deq[8 * r + i] = wrap16(I[8 * r + i] * q[8 * r + i]);
            const float s = inv_scale(r) * inv_scale(i);
            x[r][i] = (float)(D[8 * r + i] - deq[8 * r + i]) * s;
            a[r][i] = (float)W[8 * r + i] * (s * (1.0f / 255.0f));
        }
    for(int r = 0; r < 8; r++) idct8(x[r]), idct8(a[r]);
    transpose(x), transpose(a);
    for(int r = 0; r < 8; r++) idct8(x[r]), idct8(a[r]);
    for(int r = 0; r < 8; r++)
        for(int i = 0; i < 8; i++) x[r][i] *= a[r][i];
    for(int r = 0; r < 8; r++) fdct8(x[r]);
    transpose(x);
    for(int r = 0; r < 8; r++) fdct8(x[r]);
    for(int r = 0; r < 8; r++)
        for(int i = 0; i < 8; i++) {
            const float Y = x[r][i] * (fwd_scale(r) * fwd_scale(i));
            I[8 * r + i] = (int16_t)tdiv(wrap16(deq[8 * r + i] + f2i_trunc(Y)), rq[8 * r + i]);
        }
    return (int)cls;
}

// the generic class exactly as k2_generic_kernel computes it: float tables qs = q*s, q, rq per
// image; dropon side Ds = D*s and pixel-domain alpha A = IDCT2(W*s/255) (k1_lists.cu); packed
// arithmetic in the two pairings P (row r; cols 2j, 2j+1) and Q (rows 2i, 2i+1; col k);
// requant_pair() for the fp32-pipe requantisation.
void emul_generic_block_v2(int16_t *I, const int16_t *D, const int16_t *W, const uint16_t *q) {
    float A[8][8];
    F2    x[32], y[32];
    for(int v = 0; v < 8; v++)
        for(int u = 0; u < 8; u++) A[v][u] = (float)W[8 * v + u] * ((inv_scale(v) * inv_scale(u)) * (1.0f / 255.0f));
    for(int v = 0; v < 8; v++) idct8(A[v]);
    transpose(A);
    for(int v = 0; v < 8; v++) idct8(A[v]);
    transpose(A); // natural [py][px]
    for(int r = 0; r < 8; r++)
        for(int j = 0; j < 4; j++) {
            F2 Ip, qs, ds;
            float *Ipp = &Ip.x, *qsp = &qs.x, *dsp = &ds.x;
            for(int h = 0; h < 2; h++) {
                const int   i = 8 * r + 2 * j + h;
                const float s = inv_scale(r) * inv_scale(2 * j + h);
                Ipp[h] = (float)I[i];
                qsp[h] = (float)q[i] * s;
                dsp[h] = (float)D[i] * s;
            }
            x[4 * r + j] = fma2(Ip, neg2(qs), ds);
        }
    for(int j = 0; j < 4; j++) idct8p_cols_to_rowpairs(x, y, j);
    for(int i = 0; i < 4; i++) idct8p<1>(y + 8 * i);
    for(int i = 0; i < 4; i++)
        for(int k = 0; k < 8; k++) y[8 * i + k] = mul2(y[8 * i + k], f2(A[2 * i][k], A[2 * i + 1][k]));
    for(int i = 0; i < 4; i++) fdct8p_rowpairs_to_cols(y, x, i);
    for(int j = 0; j < 4; j++) fdct8p<4>(x + j);
    for(int r = 0; r < 8; r++)
        for(int j = 0; j < 4; j++) {
            const int      i = 8 * r + 2 * j;
            const F2       f = f2((float)((double)fwd_scale_d(r) * fwd_scale_d(2 * j)), (float)((double)fwd_scale_d(r) * fwd_scale_d(2 * j + 1)));
            const uint32_t pk = requant_pair(x[4 * r + j], f, f2((float)I[i], (float)I[i + 1]), f2((float)q[i], (float)q[i + 1]),
                                             f2(quant_rcp_f((float)q[i]), quant_rcp_f((float)q[i + 1])));
            I[i] = (int16_t)(pk & 0xffffu);
            I[i + 1] = (int16_t)(pk >> 16);
        }
}

// exhaustive check of the requantisation's truncating division: low int16 of requant_pair with
// y = 0 must equal a / q for a = I*q + 0 ... covered through emul_requant_check below
long long emul_requant_check(int q, int amin, int amax) {
    long long   bad = 0;
    const float rq = quant_rcp_f((float)q);
    for(int a = amin; a <= amax; a++) {
        // a = I*q + t with I = 0: feed t through y (f = 1): t = trunc(y) = a
        const uint32_t pk = requant_pair(f2((float)a + (a < 0 ? -0.25f : 0.25f), (float)a), f2(1.0f, 1.0f), f2(0.0f, 0.0f), f2((float)q, (float)q), f2(rq, rq));
        const int      want = a / q;
        if((int16_t)(pk & 0xffffu) != (int16_t)want || (int16_t)(pk >> 16) != (int16_t)want) bad++;
    }
    return bad;
}

// the fp32-pipe U / OPAQUE arithmetic of k2_fast_kernel (uniform_pair, tdiv_pair) against the integer
// formulation blend_uniform() / tdiv() that the oracle tests pin; pseudo-random sweep, returns mismatches
long long emul_uniform_pair_check(int q, int wdc, int n, unsigned seed) {
    long long   bad = 0;
    const float rq = quant_rcp(q), w4 = uniform_w4(wdc);
    unsigned    s = seed * 2654435761u + 12345u;
    for(int i = 0; i < n; i++) {
        s = s * 1664525u + 1013904223u;
        const int lim = 32767 / q;
        const int I = (int)((s >> 8) % (unsigned)(2 * lim + 1)) - lim; // |I*q| <= 32767: no int16 wrap
        s = s * 1664525u + 1013904223u;
        const int D = (int)((s >> 8) % 16601u) - 8300;
        const int want = blend_uniform(I, D, q, rq, w4);
        const uint32_t pk = uniform_pair(f2((float)I, (float)I), f2((float)D, (float)D), f2((float)q, (float)q), f2(rq, rq), w4);
        if((int16_t)(pk & 0xffffu) != (int16_t)want || (int16_t)(pk >> 16) != (int16_t)want) bad++;
        const uint32_t po = tdiv_pair(f2((float)D, (float)D), f2(rq, rq));
        if((int16_t)(po & 0xffffu) != (int16_t)(D / q)) bad++;
    }
    return bad;
}

void emul_compose_plane(int16_t *plane, int stride_blocks, int x0, int y0, const int16_t *Dp, const int16_t *Wp, int wb,
                        int hb, const uint16_t *q, long long *class_counts) {
    for(int l = 0; l < hb; l++)
        for(int k = 0; k < wb; k++) {
            int cls = emul_compose_block(plane + ((size_t)(y0 + l) * stride_blocks + x0 + k) * 64,
                                         Dp + ((size_t)l * wb + k) * 64, Wp + ((size_t)l * wb + k) * 64, q);
            if(class_counts) class_counts[cls]++;
        }
}

// islow forward DCT + q=1 quantisation of one 8x8 block of level-shifted samples (K1's core)
void emul_fdct_islow(const int *samples, int16_t *out) {
    int m[8][8];
    for(int r = 0; r < 8; r++) {
        int v[8];
        for(int i = 0; i < 8; i++) v[i] = samples[8 * r + i];
        fdct8_islow<0>(v);
        for(int i = 0; i < 8; i++) m[r][i] = v[i];
    }
    for(int c = 0; c < 8; c++) {
        int v[8];
        for(int i = 0; i < 8; i++) v[i] = m[i][c];
        fdct8_islow<1>(v);
        for(int i = 0; i < 8; i++) out[8 * i + c] = (int16_t)quant_q1(v[i]);
    }
}
}
