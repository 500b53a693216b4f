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

// libjpeg-turbo jccolor.c RGB -> YCbCr, 16-bit fixed point
MJX_HD int rgb_to_y(int r, int g, int b) { return (19595 * r + 38470 * g + 7471 * b + 32768) >> 16; }
MJX_HD int rgb_to_cb(int r, int g, int b) { return (-11059 * r - 21709 * g + 32768 * b + (128 << 16) + 32767) >> 16; }
MJX_HD int rgb_to_cr(int r, int g, int b) { return (32768 * r - 27439 * g - 5329 * b + (128 << 16) + 32767) >> 16; }

} // namespace mjx
