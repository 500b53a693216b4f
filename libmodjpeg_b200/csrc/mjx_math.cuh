// mjx_math.cuh -- per-lane arithmetic shared by the K1/K2/K3 kernels.
//
// Thread mapping used everywhere: one 8x8 coefficient block (64 int16 = 128 B, natural order,
// index 8*v + u) is owned by 8 consecutive lanes; lane r holds row r (v = r) as one 128-bit
// word = 8 int16, so a warp touches 4 consecutive blocks = 512 contiguous bytes per load.
//
// Everything here is MJX_HD so tests/host_emul can compile the same arithmetic with g++ and
// compare it with the oracle on the CPU box (test aid only -- the product has no CPU path).
#pragma once

#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define MJX_HD __host__ __device__ __forceinline__
#else
#define MJX_HD static inline
#endif

namespace mjx {

// ---------------------------------------------------------------------------------------
// block classes written by K1 into the compiled dropon's meta word
// ---------------------------------------------------------------------------------------
enum : uint32_t {
    CLS_T = 0,      // all 64 alpha coefficients zero: block untouched, never loaded
    CLS_U = 1,      // only the alpha DC is non-zero: uniform blend, exact integer/fp32 formula
    CLS_OPAQUE = 2, // uniform with alpha DC == 2040 (w0 == 0.25f): result = tdiv(D, q), image not read
    CLS_G = 3       // anything else: pixel-domain blend through three 2-D 8x8 transforms
};

MJX_HD uint32_t meta_pack(uint32_t cls, int wdc) { return cls | ((uint32_t)(uint16_t)(int16_t)wdc << 8); }
MJX_HD uint32_t meta_cls(uint32_t m) { return m & 0xffu; }
MJX_HD int meta_wdc(uint32_t m) { return (int)(int16_t)(uint16_t)(m >> 8); }

// ---------------------------------------------------------------------------------------
// 8 x int16 <-> 128-bit word
// ---------------------------------------------------------------------------------------
struct Row8 {
    uint32_t w[4];
};

MJX_HD int row_get(const Row8 &r, int i) { // i compile-time after unrolling
    uint32_t v = r.w[i >> 1];
    return (i & 1) ? ((int32_t)v >> 16) : (int)(int16_t)(uint16_t)(v & 0xffffu);
}

MJX_HD void row_unpack(const Row8 &r, int *v) {
#pragma unroll
    for(int i = 0; i < 8; i++) v[i] = row_get(r, i);
}

MJX_HD Row8 row_pack(const int *v) {
    Row8 r;
#pragma unroll
    for(int i = 0; i < 4; i++) r.w[i] = ((uint32_t)v[2 * i] & 0xffffu) | ((uint32_t)v[2 * i + 1] << 16);
    return r;
}

MJX_HD int wrap16(int x) { return (int)(int16_t)(uint16_t)(uint32_t)x; }

// ---------------------------------------------------------------------------------------
// truncating division by a quantiser value through a biased fp32 reciprocal.
//   tdiv(a, q) == a / q (C semantics) for every a in [-32768, 32767], q in [1, 65535]
// with rq = quant_rcp(q): rq = fl(fl(1/q) * (1 + 2^-21)) > (1/q)(1 + 2^-22), so exact
// multiples never fall below the integer and (k+1)q <= 98303 keeps non-multiples below k+1
// (exhaustively checked in tests/test_host_emul.py).
// ---------------------------------------------------------------------------------------
MJX_HD float quant_rcp(int q) { return (1.0f / (float)q) * 1.000000476837158203125f; }
MJX_HD float quant_rcp_f(float q) { return (1.0f / q) * 1.000000476837158203125f; } // q already converted
#if defined(__CUDACC__)
// The same through MUFU.RCP (one instruction instead of an IEEE division chain) for the per-image tables of K2:
// trunc(a * rq) == a / q holds for every rq in [1/q, (1/q)(1 + 2^-17.6)] when |a| <= 2^17; rcp.approx is within
// 2^-22 of 1/q, so rcp * (1 + 2^-20) lies inside that window.  Verified exhaustively on the device for every
// q in [1, 65535] and |a| <= 2^17 by mjx_selftest_reciprocal (tests/test_gpu_parity.py).
__device__ __forceinline__ float quant_rcp_fast(float q) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(q));
    return __fmul_rn(r, 1.00000095367431640625f);
}
#endif

MJX_HD int tdiv(int a, float rq) {
#if defined(__CUDA_ARCH__)
    return __float2int_rz(__fmul_rn((float)a, rq));
#else
    return (int)((float)a * rq);
#endif
}

// ---------------------------------------------------------------------------------------
// uniform-alpha weight (reference: src/dropon.c:548): w0 = (float)((float)W[0] * (c0*c0/1020))
// and the blend term Y = (float)(4 * (double)X * (double)w0) (reference: src/convolve.c:38,569
// with k = l = 0).  X is an integer < 2^24 and w0 a float, so the double product is exact and
// one fp32 multiply by 4*w0 rounds identically.
// ---------------------------------------------------------------------------------------
MJX_HD float uniform_w4(int wdc) {
    const double k = 0.3535534 * 0.3535534 / 1020.0;
    float w0 = (float)((double)(float)wdc * k);
    return 4.0f * w0;
}

MJX_HD float fmul_exact(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fmul_rn(a, b);
#else
    return a * b;
#endif
}

MJX_HD int f2i_trunc(float x) {
#if defined(__CUDA_ARCH__)
    return __float2int_rz(x);
#else
    return (int)x;
#endif
}

// one coefficient of a uniform-alpha block (reference: src/compose.c:277-336)
MJX_HD int blend_uniform(int I, int D, int q, float rq, float w4) {
    int deq = wrap16(I * q);
    float X = (float)(D - deq);
    int t = wrap16(deq + f2i_trunc(fmul_exact(X, w4)));
    return tdiv(t, rq);
}

// ---------------------------------------------------------------------------------------
// scaled 8-point DCTs (Arai-Agui-Nakajima flow graphs, 5 multiplies + 29 adds each).
//   forward:  true orthonormal DCT-II coefficient k = out[k] * kFwdScale[k]
//   inverse:  feed in[k] = coefficient k * kInvScale[k], get the orthonormal inverse
// with kInvScale[k] = a[k] / sqrt(8), kFwdScale[k] = 1 / (a[k] sqrt(8)),
// a[0] = 1, a[k] = sqrt(2) cos(k pi / 16).
// ---------------------------------------------------------------------------------------
#define MJX_INV_SCALE_INIT {0.35355339059327376f, 0.49039264020161522f, 0.46193976625564337f, 0.41573480615127262f, \
                            0.35355339059327376f, 0.27778511650980111f, 0.19134171618254489f, 0.09754516100806413f}
#define MJX_FWD_SCALE_INIT {0.35355339059327376f, 0.25489778955207959f, 0.27059805007309851f, 0.30067244346752264f, \
                            0.35355339059327376f, 0.44998811156820786f, 0.65328148243818826f, 1.28145772387075308f}

MJX_HD float inv_scale(int k) {
    const float t[8] = MJX_INV_SCALE_INIT;
    return t[k];
}
MJX_HD double fwd_scale_d(int k) { // the constants c_fwd2 in k2_compose.cu is built from
    const double t[8] = {0.35355339059327376, 0.25489778955207959, 0.27059805007309851, 0.30067244346752264,
                         0.35355339059327376, 0.44998811156820786, 0.65328148243818826, 1.28145772387075308};
    return t[k];
}
MJX_HD float fwd_scale(int k) {
    const float t[8] = MJX_FWD_SCALE_INIT;
    return t[k];
}

MJX_HD void idct8(float *v) {
    float t10 = v[0] + v[4], t11 = v[0] - v[4];
    float t13 = v[2] + v[6], t12 = (v[2] - v[6]) * 1.414213562373095049f - t13;
    float t0 = t10 + t13, t3 = t10 - t13, t1 = t11 + t12, t2 = t11 - t12;
    float z13 = v[5] + v[3], z10 = v[5] - v[3], z11 = v[1] + v[7], z12 = v[1] - v[7];
    float t7 = z11 + z13;
    float u11 = (z11 - z13) * 1.414213562373095049f;
    float z5 = (z10 + z12) * 1.847759065022573512f;
    float u10 = z5 - z12 * 1.082392200292393968f;
    float u12 = z5 - z10 * 2.613125929752753056f;
    float t6 = u12 - t7, t5 = u11 - t6, t4 = u10 - t5;
    v[0] = t0 + t7;
    v[7] = t0 - t7;
    v[1] = t1 + t6;
    v[6] = t1 - t6;
    v[2] = t2 + t5;
    v[5] = t2 - t5;
    v[3] = t3 + t4;
    v[4] = t3 - t4;
}

MJX_HD void fdct8(float *v) {
    float t0 = v[0] + v[7], t7 = v[0] - v[7], t1 = v[1] + v[6], t6 = v[1] - v[6];
    float t2 = v[2] + v[5], t5 = v[2] - v[5], t3 = v[3] + v[4], t4 = v[3] - v[4];
    float t10 = t0 + t3, t13 = t0 - t3, t11 = t1 + t2, t12 = t1 - t2;
    v[0] = t10 + t11;
    v[4] = t10 - t11;
    float z1 = (t12 + t13) * 0.707106781186547524f;
    v[2] = t13 + z1;
    v[6] = t13 - z1;
    t10 = t4 + t5;
    t11 = t5 + t6;
    t12 = t6 + t7;
    float z5 = (t10 - t12) * 0.382683432365089772f;
    float z2 = 0.541196100146196985f * t10 + z5;
    float z4 = 1.306562964876376527f * t12 + z5;
    float z3 = t11 * 0.707106781186547524f;
    float z11 = t7 + z3, z13 = t7 - z3;
    v[5] = z13 + z2;
    v[3] = z13 - z2;
    v[1] = z11 + z4;
    v[7] = z11 - z4;
}

// ---------------------------------------------------------------------------------------
// libjpeg-turbo's integer forward DCT (jfdctint.c jpeg_fdct_islow, CONST_BITS 13,
// PASS1_BITS 2) -- K1 must reproduce it bit for bit.  PASS 0 = rows, 1 = columns.
// ---------------------------------------------------------------------------------------
template <int PASS>
MJX_HD void fdct8_islow(int *d) {
    const int SH = PASS == 0 ? 11 : 15;
    const int RND = 1 << (SH - 1);
    int tmp0 = d[0] + d[7], tmp7 = d[0] - d[7], tmp1 = d[1] + d[6], tmp6 = d[1] - d[6];
    int tmp2 = d[2] + d[5], tmp5 = d[2] - d[5], tmp3 = d[3] + d[4], tmp4 = d[3] - d[4];
    int tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
    if(PASS == 0) {
        d[0] = (tmp10 + tmp11) * 4;
        d[4] = (tmp10 - tmp11) * 4;
    }
    else {
        d[0] = (tmp10 + tmp11 + 2) >> 2;
        d[4] = (tmp10 - tmp11 + 2) >> 2;
    }
    int z1 = (tmp12 + tmp13) * 4433;
    d[2] = (z1 + tmp13 * 6270 + RND) >> SH;
    d[6] = (z1 - tmp12 * 15137 + RND) >> SH;
    z1 = tmp4 + tmp7;
    int z2 = tmp5 + tmp6, z3 = tmp4 + tmp6, z4 = tmp5 + tmp7;
    int z5 = (z3 + z4) * 9633;
    tmp4 *= 2446;
    tmp5 *= 16819;
    tmp6 *= 25172;
    tmp7 *= 12299;
    z1 *= -7373;
    z2 *= -20995;
    z3 = z3 * -16069 + z5;
    z4 = z4 * -3196 + z5;
    d[7] = (tmp4 + z1 + z3 + RND) >> SH;
    d[5] = (tmp5 + z2 + z4 + RND) >> SH;
    d[3] = (tmp6 + z2 + z3 + RND) >> SH;
    d[1] = (tmp7 + z1 + z4 + RND) >> SH;
}

// quality-100 quantisation of an islow output (divisor 8): sign(x) * ((|x| + 4) >> 3)
MJX_HD int quant_q1(int x) { return x < 0 ? -((-x + 4) >> 3) : ((x + 4) >> 3); }

// ---------------------------------------------------------------------------------------
// whole-block (thread-per-block) arithmetic of the generic class.  The same AAN flow graphs as
// idct8/fdct8 but on strided register arrays so a thread can transform rows (S = 1) and
// columns (S = 8) of its 64 registers without any transpose, written so that multiply+add
// pairs contract to FFMA.  No int<->float conversion instructions are used on the hot path:
// on sm_100a F2I/FRND issue at 1/8 rate and I2F.S16 / SHFL at 1/4 (profiles/microbench).
// ---------------------------------------------------------------------------------------
template <int S>
MJX_HD void idct8s(float *v) {
    float t10 = v[0] + v[4 * S], t11 = v[0] - v[4 * S];
    float t13 = v[2 * S] + v[6 * S], t12 = (v[2 * S] - v[6 * S]) * 1.414213562373095049f - t13;
    float t0 = t10 + t13, t3 = t10 - t13, t1 = t11 + t12, t2 = t11 - t12;
    float z13 = v[5 * S] + v[3 * S], z10 = v[5 * S] - v[3 * S], z11 = v[S] + v[7 * S], z12 = v[S] - v[7 * S];
    float t7 = z11 + z13;
    float z5 = (z10 + z12) * 1.847759065022573512f;
    float t6 = (z5 - z10 * 2.613125929752753056f) - t7;
    float t5 = (z11 - z13) * 1.414213562373095049f - t6;
    float t4 = (z5 - z12 * 1.082392200292393968f) - t5;
    v[0] = t0 + t7;
    v[7 * S] = t0 - t7;
    v[S] = t1 + t6;
    v[6 * S] = t1 - t6;
    v[2 * S] = t2 + t5;
    v[5 * S] = t2 - t5;
    v[3 * S] = t3 + t4;
    v[4 * S] = t3 - t4;
}

template <int S>
MJX_HD void fdct8s(float *v) {
    float t0 = v[0] + v[7 * S], t7 = v[0] - v[7 * S], t1 = v[S] + v[6 * S], t6 = v[S] - v[6 * S];
    float t2 = v[2 * S] + v[5 * S], t5 = v[2 * S] - v[5 * S], t3 = v[3 * S] + v[4 * S], t4 = v[3 * S] - v[4 * S];
    float t10 = t0 + t3, t13 = t0 - t3, t11 = t1 + t2, t12 = t1 - t2;
    v[0] = t10 + t11;
    v[4 * S] = t10 - t11;
    float z1 = (t12 + t13) * 0.707106781186547524f;
    v[2 * S] = t13 + z1;
    v[6 * S] = t13 - z1;
    t10 = t4 + t5;
    t11 = t5 + t6;
    t12 = t6 + t7;
    float z5 = (t10 - t12) * 0.382683432365089772f;
    float z2 = 0.541196100146196985f * t10 + z5;
    float z4 = 1.306562964876376527f * t12 + z5;
    float z11 = t7 + t11 * 0.707106781186547524f, z13 = t7 - t11 * 0.707106781186547524f;
    v[5 * S] = z13 + z2;
    v[3 * S] = z13 - z2;
    v[S] = z11 + z4;
    v[7 * S] = z11 - z4;
}

// truncation toward zero that stays in the fp32 pipe: |v| + 2^23 rounded toward zero has ulp 1,
// so it is 2^23 + floor(|v|); exact for |v| < 2^23.  (== (float)(int)v, the reference's (int)Y.)
MJX_HD float trunc_f(float v) {
#if defined(__CUDA_ARCH__)
    const float m = __fadd_rz(fabsf(v), 8388608.0f) - 8388608.0f;
    return copysignf(m, v);
#else
    return (float)(int)v;
#endif
}

// low 16 bits of (integer-valued float o + 1.5*2^23) are o as a two's-complement int16
MJX_HD uint32_t int16_bits_of(float o) {
    const float  biased = o + 12582912.0f;
    uint32_t     u;
#if defined(__CUDA_ARCH__)
    u = __float_as_uint(biased);
#else
    memcpy(&u, &biased, 4);
#endif
    return u;
}

MJX_HD uint32_t pack2_int16(float lo, float hi) {
#if defined(__CUDA_ARCH__)
    return __byte_perm(int16_bits_of(lo), int16_bits_of(hi), 0x5410);
#else
    return (int16_bits_of(lo) & 0xffffu) | (int16_bits_of(hi) << 16);
#endif
}

// requantisation of one generic coefficient in the fp32 pipe: (deq + (int)Y) / q, truncating.
// Equal to tdiv(deq + (int)Y, rq) whenever no int16 wrap-around occurs (|deq + Y| < 32768),
// which holds for every JPEG a conforming encoder produces; the strict kernel keeps the wraps.
MJX_HD float requant_f(float deq_f, float Y, float rq) { return trunc_f((deq_f + trunc_f(Y)) * rq); }

// ---------------------------------------------------------------------------------------
// packed fp32 (sm_100a FADD2 / FMUL2 / FFMA2: two fp32 lanes per instruction, per-operand negate
// and a broadcast scalar immediate are free).  The generic class runs entirely on these: one
// thread owns one block as 32 pairs, so every 1-D pass is 30 packed instructions per pair of
// rows / columns instead of 60 scalar ones.  Host versions (test aid) round identically.
// ---------------------------------------------------------------------------------------
#if defined(__CUDACC__)
typedef float2 F2;
#else
struct F2 {
    float x, y;
};
#endif
MJX_HD F2 f2(float x, float y) {
    F2 r;
    r.x = x, r.y = y;
    return r;
}
MJX_HD F2 neg2(F2 a) { return f2(-a.x, -a.y); }
// a*b + c rounded toward zero in one step (host model): the product of two floats is exact in
// double, and here |a*b| < 2^23 = |c| with equal signs, so truncating the product and adding c
// lands on the float grid (spacing 1) exactly where the hardware's single RZ rounding does
MJX_HD float fma_rz_magic(float a, float b, float c) {
    const double p = (double)a * (double)b;
    const double t = p < 0 ? ceil(p) : floor(p);
    return (float)(t + (double)c);
}
#if defined(__CUDA_ARCH__)
MJX_HD F2 add2(F2 a, F2 b) { return __fadd2_rn(a, b); }
MJX_HD F2 sub2(F2 a, F2 b) { return __fadd2_rn(a, neg2(b)); }
MJX_HD F2 mul2(F2 a, F2 b) { return __fmul2_rn(a, b); }
MJX_HD F2 fma2(F2 a, F2 b, F2 c) { return __ffma2_rn(a, b, c); }
MJX_HD F2 fma2_rz(F2 a, F2 b, F2 c) { return __ffma2_rz(a, b, c); }
MJX_HD F2 add2_rz(F2 a, F2 b) { return __fadd2_rz(a, b); }
MJX_HD float add1(float a, float b) { return __fadd_rn(a, b); }
MJX_HD float sub1(float a, float b) { return __fadd_rn(a, -b); }
#else
MJX_HD F2 add2(F2 a, F2 b) { return f2(a.x + b.x, a.y + b.y); }
MJX_HD F2 sub2(F2 a, F2 b) { return f2(a.x - b.x, a.y - b.y); }
MJX_HD F2 mul2(F2 a, F2 b) { return f2(a.x * b.x, a.y * b.y); }
MJX_HD F2 fma2(F2 a, F2 b, F2 c) { return f2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)); }
MJX_HD F2 fma2_rz(F2 a, F2 b, F2 c) { return f2(fma_rz_magic(a.x, b.x, c.x), fma_rz_magic(a.y, b.y, c.y)); }
MJX_HD F2 add2_rz(F2 a, F2 b) { return f2(fma_rz_magic(a.x, 1.0f, b.x), fma_rz_magic(a.y, 1.0f, b.y)); } // b = +-2^23, sign of a
MJX_HD float add1(float a, float b) { return a + b; }
MJX_HD float sub1(float a, float b) { return a - b; }
#endif
MJX_HD F2 bc2(float c) { return f2(c, c); }

// 8-point AAN inverse DCT on pairs, elements v[0], v[S], .. v[7S].  If TO is non-null the last
// butterfly stage is done with scalar adds that write the result in the OTHER pairing (see
// k2_generic_kernel): the pair (.x, .y) of output element k goes to TO[..] halves chosen by the
// caller through the two index maps below.  Plain form first.
template <int S>
MJX_HD void idct8p(F2 *v) {
    F2 t10 = add2(v[0], v[4 * S]), t11 = sub2(v[0], v[4 * S]);
    F2 t13 = add2(v[2 * S], v[6 * S]), t12 = fma2(sub2(v[2 * S], v[6 * S]), bc2(1.414213562373095049f), neg2(t13));
    F2 t0 = add2(t10, t13), t3 = sub2(t10, t13), t1 = add2(t11, t12), t2 = sub2(t11, t12);
    F2 z13 = add2(v[5 * S], v[3 * S]), z10 = sub2(v[5 * S], v[3 * S]), z11 = add2(v[S], v[7 * S]), z12 = sub2(v[S], v[7 * S]);
    F2 t7 = add2(z11, z13);
    F2 z5 = mul2(add2(z10, z12), bc2(1.847759065022573512f));
    F2 t6 = sub2(fma2(z10, bc2(-2.613125929752753056f), z5), t7);
    F2 t5 = fma2(sub2(z11, z13), bc2(1.414213562373095049f), neg2(t6));
    F2 t4 = sub2(fma2(z12, bc2(-1.082392200292393968f), z5), t5);
    v[0] = add2(t0, t7);
    v[7 * S] = sub2(t0, t7);
    v[S] = add2(t1, t6);
    v[6 * S] = sub2(t1, t6);
    v[2 * S] = add2(t2, t5);
    v[5 * S] = sub2(t2, t5);
    v[3 * S] = add2(t3, t4);
    v[4 * S] = sub2(t3, t4);
}

// the same, for column pair j of a block held as P pairs  x[4r + j] = (row r, cols 2j, 2j+1);
// the result is written as Q pairs  y[8i + k] = (rows 2i, 2i+1; col k)  by a scalar last stage
MJX_HD void idct8p_cols_to_rowpairs(const F2 *x, F2 *y, int j) {
    const F2 *v = x + j;
    F2 t10 = add2(v[0], v[16]), t11 = sub2(v[0], v[16]);
    F2 t13 = add2(v[8], v[24]), t12 = fma2(sub2(v[8], v[24]), bc2(1.414213562373095049f), neg2(t13));
    F2 t0 = add2(t10, t13), t3 = sub2(t10, t13), t1 = add2(t11, t12), t2 = sub2(t11, t12);
    F2 z13 = add2(v[20], v[12]), z10 = sub2(v[20], v[12]), z11 = add2(v[4], v[28]), z12 = sub2(v[4], v[28]);
    F2 t7 = add2(z11, z13);
    F2 z5 = mul2(add2(z10, z12), bc2(1.847759065022573512f));
    F2 t6 = sub2(fma2(z10, bc2(-2.613125929752753056f), z5), t7);
    F2 t5 = fma2(sub2(z11, z13), bc2(1.414213562373095049f), neg2(t6));
    F2 t4 = sub2(fma2(z12, bc2(-1.082392200292393968f), z5), t5);
    // rows 0..7 = t0+t7, t1+t6, t2+t5, t3+t4, t3-t4, t2-t5, t1-t6, t0-t7
    y[0 + 2 * j] = f2(add1(t0.x, t7.x), add1(t1.x, t6.x));
    y[0 + 2 * j + 1] = f2(add1(t0.y, t7.y), add1(t1.y, t6.y));
    y[8 + 2 * j] = f2(add1(t2.x, t5.x), add1(t3.x, t4.x));
    y[8 + 2 * j + 1] = f2(add1(t2.y, t5.y), add1(t3.y, t4.y));
    y[16 + 2 * j] = f2(sub1(t3.x, t4.x), sub1(t2.x, t5.x));
    y[16 + 2 * j + 1] = f2(sub1(t3.y, t4.y), sub1(t2.y, t5.y));
    y[24 + 2 * j] = f2(sub1(t1.x, t6.x), sub1(t0.x, t7.x));
    y[24 + 2 * j + 1] = f2(sub1(t1.y, t6.y), sub1(t0.y, t7.y));
}

template <int S>
MJX_HD void fdct8p(F2 *v) {
    F2 t0 = add2(v[0], v[7 * S]), t7 = sub2(v[0], v[7 * S]), t1 = add2(v[S], v[6 * S]), t6 = sub2(v[S], v[6 * S]);
    F2 t2 = add2(v[2 * S], v[5 * S]), t5 = sub2(v[2 * S], v[5 * S]), t3 = add2(v[3 * S], v[4 * S]), t4 = sub2(v[3 * S], v[4 * S]);
    F2 t10 = add2(t0, t3), t13 = sub2(t0, t3), t11 = add2(t1, t2), t12 = sub2(t1, t2);
    v[0] = add2(t10, t11);
    v[4 * S] = sub2(t10, t11);
    F2 z1 = mul2(add2(t12, t13), bc2(0.707106781186547524f));
    v[2 * S] = add2(t13, z1);
    v[6 * S] = sub2(t13, z1);
    t10 = add2(t4, t5);
    t11 = add2(t5, t6);
    t12 = add2(t6, t7);
    F2 z5 = mul2(sub2(t10, t12), bc2(0.382683432365089772f));
    F2 z2 = fma2(t10, bc2(0.541196100146196985f), z5);
    F2 z4 = fma2(t12, bc2(1.306562964876376527f), z5);
    F2 z11 = fma2(t11, bc2(0.707106781186547524f), t7), z13 = fma2(t11, bc2(-0.707106781186547524f), t7);
    v[5 * S] = add2(z13, z2);
    v[3 * S] = sub2(z13, z2);
    v[S] = add2(z11, z4);
    v[7 * S] = sub2(z11, z4);
}

// forward transform of row pair i held as Q pairs y[8i + k]; the result is written as P pairs
// x[4r + j] = (row r, cols 2j, 2j+1) by a scalar last stage
MJX_HD void fdct8p_rowpairs_to_cols(const F2 *y, F2 *x, int i) {
    const F2 *v = y + 8 * i;
    F2 t0 = add2(v[0], v[7]), t7 = sub2(v[0], v[7]), t1 = add2(v[1], v[6]), t6 = sub2(v[1], v[6]);
    F2 t2 = add2(v[2], v[5]), t5 = sub2(v[2], v[5]), t3 = add2(v[3], v[4]), t4 = sub2(v[3], v[4]);
    F2 t10 = add2(t0, t3), t13 = sub2(t0, t3), t11 = add2(t1, t2), t12 = sub2(t1, t2);
    F2 z1 = mul2(add2(t12, t13), bc2(0.707106781186547524f));
    F2 u10 = add2(t4, t5), u11 = add2(t5, t6), u12 = add2(t6, t7);
    F2 z5 = mul2(sub2(u10, u12), bc2(0.382683432365089772f));
    F2 z2 = fma2(u10, bc2(0.541196100146196985f), z5);
    F2 z4 = fma2(u12, bc2(1.306562964876376527f), z5);
    F2 z11 = fma2(u11, bc2(0.707106781186547524f), t7), z13 = fma2(u11, bc2(-0.707106781186547524f), t7);
    // cols 0..7 = t10+t11, z11+z4, t13+z1, z13-z2, t10-t11, z13+z2, t13-z1, z11-z4
    F2 *lo = x + 4 * (2 * i), *hi = x + 4 * (2 * i + 1);
    lo[0] = f2(add1(t10.x, t11.x), add1(z11.x, z4.x));
    hi[0] = f2(add1(t10.y, t11.y), add1(z11.y, z4.y));
    lo[1] = f2(add1(t13.x, z1.x), sub1(z13.x, z2.x));
    hi[1] = f2(add1(t13.y, z1.y), sub1(z13.y, z2.y));
    lo[2] = f2(sub1(t10.x, t11.x), add1(z13.x, z2.x));
    hi[2] = f2(sub1(t10.y, t11.y), add1(z13.y, z2.y));
    lo[3] = f2(sub1(t13.x, z1.x), sub1(z11.x, z4.x));
    hi[3] = f2(sub1(t13.y, z1.y), sub1(z11.y, z4.y));
}

// +-2^23 with the sign of v: the magic addend that makes a round-toward-zero add truncate
MJX_HD float signed_magic(float v) {
    uint32_t u;
#if defined(__CUDA_ARCH__)
    u = __float_as_uint(v);
#else
    memcpy(&u, &v, 4);
#endif
    u = (u & 0x80000000u) | 0x4B000000u;
#if defined(__CUDA_ARCH__)
    return __uint_as_float(u);
#else
    float r;
    memcpy(&r, &u, 4);
    return r;
#endif
}
MJX_HD F2 signed_magic2(F2 v) { return f2(signed_magic(v.x), signed_magic(v.y)); }

// one pair of generic coefficients, fp32 pipe only:
//   t   = trunc(y * f)                       (the reference's (int)Y, src/compose.c:315-324)
//   a   = I*q + t                            (dequantised + blend term; exact, integers < 2^24)
//   out = trunc(a / q) as int16 bits         (src/compose.c:327-336)
// trunc(y*f) comes from ONE round-toward-zero FMA onto +-2^23, trunc(a/q) from one onto the biased
// reciprocal (quant_rcp, exact for |a| < 2^20, tests/test_host_emul.py); the int16 bit pattern is
// the low half of (value + 1.5*2^23).  Returns the two int16 packed in one word.
MJX_HD uint32_t requant_pair(F2 y, F2 f, F2 I, F2 q, F2 rq) {
    const F2 sm = signed_magic2(y);
    const F2 t = sub2(fma2_rz(y, f, sm), sm);
    const F2 a = fma2(I, q, t);
    const F2 sa = signed_magic2(a);
    const F2 m = fma2_rz(a, rq, sa);
    const F2 o = add2(m, sub2(bc2(12582912.0f), sa));
    uint32_t lo, hi;
#if defined(__CUDA_ARCH__)
    lo = __float_as_uint(o.x), hi = __float_as_uint(o.y);
    return __byte_perm(lo, hi, 0x5410);
#else
    memcpy(&lo, &o.x, 4), memcpy(&hi, &o.y, 4);
    return (lo & 0xffffu) | (hi << 16);
#endif
}

// trunc(a / q) of a pair as two packed int16 (second half of requant_pair)
MJX_HD uint32_t tdiv_pair(F2 a, F2 rq) {
    const F2 sa = signed_magic2(a);
    const F2 m = fma2_rz(a, rq, sa);
    const F2 o = add2(m, sub2(bc2(12582912.0f), sa));
    uint32_t lo, hi;
#if defined(__CUDA_ARCH__)
    lo = __float_as_uint(o.x), hi = __float_as_uint(o.y);
    return __byte_perm(lo, hi, 0x5410);
#else
    memcpy(&lo, &o.x, 4), memcpy(&hi, &o.y, 4);
    return (lo & 0xffffu) | (hi << 16);
#endif
}

// one pair of uniform-alpha coefficients in the fp32 pipe, bit-exact with the reference whenever
// no int16 wrap-around occurs (src/compose.c:277-336 with the single non-zero weight w0):
//   Iq = I*q (exact), X = D - Iq (exact), Y = fl(X * w4)  [== (float)(4 * (double)X * (double)w0)],
//   a = Iq + trunc(Y), out = trunc(a / q)
MJX_HD uint32_t uniform_pair(F2 I, F2 D, F2 q, F2 rq, float w4) {
    const F2 Iq = mul2(I, q);
    const F2 Y = mul2(sub2(D, Iq), bc2(w4));
    const F2 sm = signed_magic2(Y);
    const F2 t = sub2(add2_rz(Y, sm), sm);
    return tdiv_pair(add2(Iq, t), rq);
}

// libjpeg-turbo jccolor.c RGB -> YCbCr, 16-bit fixed point
MJX_HD int rgb_to_y(int r, int g, int b) { return (19595 * r + 38470 * g + 7471 * b + 32768) >> 16; }
MJX_HD int rgb_to_cb(int r, int g, int b) { return (-11059 * r - 21709 * g + 32768 * b + (128 << 16) + 32767) >> 16; }
MJX_HD int rgb_to_cr(int r, int g, int b) { return (32768 * r - 27439 * g - 5329 * b + (128 << 16) + 32767) >> 16; }

} // namespace mjx
