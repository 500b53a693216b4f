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

// the generic class as the thread-per-block kernel computes it (k2 generic kernel): float tables
// qs = q*s, qf = q, rq; dropon side Ds = D*s and pixel-domain alpha A = IDCT2(W*s/255); fp32-pipe
// requantisation without int<->float conversions.
void emul_generic_block_v2(int16_t *I, const int16_t *D, const int16_t *W, const uint16_t *q) {
    float x[64], A[64], deq[64];
    for(int v = 0; v < 8; v++)
        for(int u = 0; u < 8; u++) {
            const int   i = 8 * v + u;
            const float s = inv_scale(v) * inv_scale(u);
            const float Ds = (float)D[i] * s, qs = (float)q[i] * s, qf = (float)q[i];
            A[i] = (float)W[i] * (s * (1.0f / 255.0f));
            deq[i] = (float)I[i] * qf;
            x[i] = Ds - (float)I[i] * qs;
        }
    for(int v = 0; v < 8; v++) idct8s<1>(A + 8 * v), idct8s<1>(x + 8 * v);
    for(int u = 0; u < 8; u++) idct8s<8>(A + u), idct8s<8>(x + u);
    for(int i = 0; i < 64; i++) x[i] *= A[i];
    for(int u = 0; u < 8; u++) fdct8s<8>(x + u);
    for(int v = 0; v < 8; v++) fdct8s<1>(x + 8 * v);
    for(int v = 0; v < 8; v++)
        for(int u = 0; u < 8; u += 2) {
            const int   i = 8 * v + u;
            const float o0 = requant_f(deq[i], x[i] * (fwd_scale(v) * fwd_scale(u)), quant_rcp(q[i]));
            const float o1 = requant_f(deq[i + 1], x[i + 1] * (fwd_scale(v) * fwd_scale(u + 1)), quant_rcp(q[i + 1]));
            const uint32_t pk = pack2_int16(o0, o1);
            I[i] = (int16_t)(pk & 0xffffu);
            I[i + 1] = (int16_t)(pk >> 16);
        }
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
