// k2_generic_op.cu -- K2 for class G blocks of LARGE BATCHES: the whole per-block blend as one tensor-core product.
// Replaces mj_compose_with_mask + mj_convolve for blocks with a non-uniform mask, like k2_generic_kernel
// (reference: src/compose.c:237-342, src/convolve.c:29-1099).
//
// What the reference computes per block is LINEAR in the dequantised image block (src/compose.c:289-312):
//     Y = sum_{k,l} w[k,l] * M_k X M_l,   X = D - I o q      =>      Y = L(D) - L(I o q)
// with L the 64 x 64 operator T[8a+b][8m+n] = sum_{k,l} w[8k+l] M_k[m][a] M_l[n][b] that depends on the dropon block's
// alpha only (w: src/dropon.c:548-566; M_k: the eight sparse 8 x 8 matrices src/convolve.c spells out).  A batch stamps ONE
// dropon on many images, so for a dropon block b the images' blocks at that position form a [images x 64] matrix of
// EXACT small integers I, and
//     Y[image][:] = K_b - I[image][:] . (diag(q) T_b)
// is a [128 images x 64] x [64 x 64] product per 128 images: tcgen05.mma, kind::f16, fp32 accumulation in tensor memory.
//   * A operand: the raw int16 rows, gathered by cp.async straight into the K-major SWIZZLE_128B layout and converted in
//     place to fp16 (I / 512, exact for I in [-1024, 1023]: one LOP3 + one HFMA2 per coefficient pair).
//   * B operand: -S diag(q) T_b, built ONCE per (compiled dropon, quantisation tables) in double precision from the
//     reference's own constants (k2_op_build_kernel), split into kPieces fp16 pieces (11 bits each) and stored in global
//     memory as ready-made shared-memory images; the bulk-copy engine brings the next block's pieces in while the current
//     block's images are processed.  64 more columns hold diag(q): the accumulator then also carries I*q / 512 exactly,
//     which is what the requantisation needs -- no int16 -> fp32 conversion anywhere.
//   * each thread reads the 64 + 64 accumulator values of ITS image from its tensor-memory lane and requantises:
//     out = trunc((I*q + trunc(K + acc)) / q), packed fp32 (mjx_math.cuh), written over the staged row, coalesced stores.
// No inverse transform, no blend, no forward transform on the CUDA cores: per block ~12 instructions per coefficient pair
// instead of ~48, which moves the class from the fp32 pipe's roofline to the HBM roofline.
//
// Exact-arithmetic parity: T is the reference's operator evaluated in double (including its 0.3535534 constant), so the
// only differences to the reference are roundings of its own fp32 accumulation -- +-1 quantisation step, rarer than with
// the closed-form fp32 kernel (tests/test_gpu_tensor_core.py).
//
// Contract: coefficients in [-1024, 1023] (what ITU-T T.81 lets an 8-bit baseline JPEG carry), quantiser values <= 255,
// tables equal to the first image's.  Everything outside it is detected here (kCheck / k2_op_prepare_kernel) and left to
// the fp32 kernel through the redo mask; nothing is assumed.
#include <cuda_fp16.h>

#include "k2_common.cuh"
#include "k2_umma.cuh"

namespace mjx {

static constexpr int kOpStageBytes = 128 * 128; // 128 image rows x 64 int16 / fp16
static constexpr int kOpStages = 3;
static constexpr int kOpPieceBytes = 64 * 128;  // 64 output rows x 64 fp16 (K-major)
static constexpr int kOpKRing = 4;
static constexpr int kOpThreads = (kOpGroups * 5 + 1) * 32; // compute warps + one MMA warp per group + load warp

// ---------------------------------------------------------------------------------------------------------------------
// per launch: which images can take the tensor-core path, and a compact address table for them
// ---------------------------------------------------------------------------------------------------------------------
// table[c * n + i] = {plane address (2 words), row stride in blocks, rows} of component c of image i, address 0 when the
// image's table differs from image 0's (or the component is not served): those (image, component) pairs are handed to the
// fp32 kernel through the redo mask.  Block 0 also compares image 0's tables with the tables the cached operator was built
// for and raises the rebuild flags.
__global__ void __launch_bounds__(128) k2_op_prepare_kernel(const OpParams p) {
    __shared__ uint32_t s_q[MJX_MAX_COMPONENTS][32];
    __shared__ int      s_ok[MJX_MAX_COMPONENTS];
    const int           ncomp = p.drop.ncomp;
    // warp c looks at component c of image 0 (the four chains of dependent loads run side by side: this launch sits in front of
    // the operator kernel on every call)
    {
        const int c = threadIdx.x >> 5, lane = threadIdx.x & 31;
        if(c < ncomp) {
            const uint32_t w = reinterpret_cast<const uint32_t *>(&p.items[0].q[c][0])[lane];
            s_q[c][lane] = w;
            unsigned qm = max(w & 0xffffu, w >> 16), qn = min(w & 0xffffu, w >> 16);
            qm = __reduce_max_sync(0xffffffffu, qm);
            qn = __reduce_min_sync(0xffffffffu, qn);
            // scale S = 2^sh: the largest power of two <= 512 that keeps |S q T| inside fp16 (|T| <= max alpha < 1.01)
            int sh = 9;
            while(sh > 0 && 1.01f * (float)qm * (float)(1 << sh) > 65000.0f) sh--;
            const int ok = (qm <= 255u && qn >= 1u) ? sh : -1;
            if(lane == 0) s_ok[c] = ok;
            if(blockIdx.x == 0) {
                const uint32_t old = reinterpret_cast<const uint32_t *>(p.op.key + c * 64)[lane];
                const bool     differ = __any_sync(0xffffffffu, old != w);
                reinterpret_cast<uint32_t *>(p.op.key + c * 64)[lane] = w;
                if(lane == 0) {
                    p.op.rebuild[c] = differ ? 1 : 0;
                    p.op.info[c] = ok;
                }
            }
        }
    }
    __syncthreads();
    // one thread per (component, image)
    const int idx = blockIdx.x * 128 + threadIdx.x;
    if(idx >= p.n * ncomp) return;
    const int               c = idx / p.n, i = idx - c * p.n;
    const mjx_image_desc_t &im = p.items[i];
    const int               ntiles = p.drop.n_generic >> 5;
    const int               t0 = p.drop.gtile_start[c], t1 = c + 1 < ncomp ? p.drop.gtile_start[c + 1] : ntiles;
    bool                    same = s_ok[c] >= 0;
    if(same && i > 0) {
        const uint4 *q = reinterpret_cast<const uint4 *>(&im.q[c][0]);
        uint4        a[8];
#pragma unroll
        for(int k = 0; k < 8; k++) a[k] = __ldg(q + k);
#pragma unroll
        for(int k = 0; k < 8; k++)
            same = same && a[k].x == s_q[c][4 * k] && a[k].y == s_q[c][4 * k + 1] && a[k].z == s_q[c][4 * k + 2] && a[k].w == s_q[c][4 * k + 3];
    }
    const unsigned long long plane = im.plane[c];
    uint4                    t = make_uint4(0u, 0u, 0u, 0u);
    if(same && plane != 0) t = make_uint4((uint32_t)plane, (uint32_t)(plane >> 32), (uint32_t)im.stride_blocks[c], (uint32_t)im.rows[c]);
    p.table[(size_t)c * p.n + i] = t;
    if(!same && plane != 0 && t1 > t0) {
        for(int tl = t0; tl < t1; tl++) p.redo_mask[(size_t)tl * p.n + i] = 0xffffffffu;
        atomicAdd(p.redo_count, 1u);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// once per (compiled dropon, quantisation tables): the operator of every G block, in double precision
// ---------------------------------------------------------------------------------------------------------------------
// M_l[j][i] of src/convolve.c (row j = output index, column i = input index), see the header comment of oracle/mj_oracle.c
// for the derivation: M_0 = 2 I; for l >= 1 column 0 has sqrt2 at row l, column i >= 1 has +1 at row |i - l| (sqrt2 when
// that row is 0), +1 at row i + l if i + l < 8, -1 at row 16 - i - l if i + l > 8.
__device__ __forceinline__ double conv_m(int l, int j, int i) {
    const double s2 = 1.41421356237309504880;
    if(l == 0) return i == j ? 2.0 : 0.0;
    if(i == 0) return j == l ? s2 : 0.0;
    double    v = 0.0;
    const int r = i > l ? i - l : l - i;
    if(j == r) v += r == 0 ? s2 : 1.0;
    if(i + l < 8 && j == i + l) v += 1.0;
    if(i + l > 8 && j == 16 - i - l) v -= 1.0;
    return v;
}

// byte offset of element (row, k) of a K-major SWIZZLE_128B tile (64 fp16 per row)
__host__ __device__ __forceinline__ int sw128_off(int row, int k) {
    return (row >> 3) * 1024 + (row & 7) * 128 + ((((k >> 3) ^ (row & 7))) << 4) + (k & 7) * 2;
}

template <int NP>
__global__ void __launch_bounds__(256) k2_op_build_kernel(const OpParams p) {
    __shared__ double s_m[8][8][8]; // [l][j][i]
    __shared__ double s_w[64], s_d[64], s_q[64];
    __shared__ double s_u[8][8][8]; // U[l][m][a] = sum_k w[8k + l] M_k[m][a]
    __shared__ double s_p[8][8][8]; // P[l][n][a] = sum_b M_l[n][b] D[8a + b]
    const int t = threadIdx.x;
    int       any = 0;
    for(int c = 0; c < p.drop.ncomp; c++) any |= p.op.rebuild[c];
    if(!any) return; // the usual launch: the tables are the ones the cache was built for
    for(int x = t; x < 512; x += 256) s_m[x >> 6][(x >> 3) & 7][x & 7] = conv_m(x >> 6, (x >> 3) & 7, x & 7);
    for(int slot = blockIdx.x; slot < p.drop.n_generic; slot += gridDim.x) {
    __syncthreads(); // the previous slot's shared arrays are no longer read
    const uint32_t e = __ldg(p.drop.list_generic + slot);
    if(e == 0xffffffffu) continue;
    const int c = entry_comp(e);
    if(!p.op.rebuild[c]) continue;
    const int sh = p.op.info[c];
    if(sh < 0) continue;
    const DropComp &dc = p.drop.comp[c];
    const size_t    bi = (size_t)entry_row(e) * dc.wb + entry_col(e);
    if(t < 64) {
        // alpha weights exactly as the reference stores them: (float)((double)(float)coef * c(v) c(u) / 1020), src/dropon.c:548-566
        const double c0 = 0.3535534, c1 = 0.5;
        const double k = ((t >> 3) == 0 ? c0 : c1) * ((t & 7) == 0 ? c0 : c1) / 1020.0;
        s_w[t] = (double)(float)((double)(float)dc.W[bi * 64 + t] * k);
        s_d[t] = (double)dc.D[bi * 64 + t];
        s_q[t] = (double)p.items[0].q[c][t];
    }
    __syncthreads();
    for(int x = t; x < 512; x += 256) {
        const int l = x >> 6, m = (x >> 3) & 7, a = x & 7;
        double    u = 0.0, pp = 0.0;
#pragma unroll
        for(int k = 0; k < 8; k++) {
            u += s_w[8 * k + l] * s_m[k][m][a];
            pp += s_m[l][m][k] * s_d[8 * a + k]; // P[l][n = m][a], b = k
        }
        s_u[l][m][a] = u;
        s_p[l][m][a] = pp;
    }
    __syncthreads();
    // B[out = 8m + n][in = 8a + b] = -S q[in] T[in][out],  T[8a+b][8m+n] = sum_l U[l][m][a] M_l[n][b]
    unsigned char *Bs = p.op.B + (size_t)slot * NP * kOpPieceBytes;
    const double   S = (double)(1 << sh);
    for(int x = t; x < 512; x += 256) {
        const int out = x >> 3, a = x & 7, m = out >> 3, n = out & 7;
        double    v[8];
#pragma unroll
        for(int b = 0; b < 8; b++) v[b] = 0.0;
#pragma unroll
        for(int l = 0; l < 8; l++) {
            const double u = s_u[l][m][a];
#pragma unroll
            for(int b = 0; b < 8; b++) v[b] += u * s_m[l][n][b];
        }
#pragma unroll
        for(int b = 0; b < 8; b++) v[b] *= -S * s_q[8 * a + b];
        const int off = sw128_off(out, 8 * a);
#pragma unroll
        for(int pc = 0; pc < NP; pc++) {
            uint32_t w[4];
#pragma unroll
            for(int b2 = 0; b2 < 4; b2++) {
                const __half h0 = __float2half_rn((float)v[2 * b2]), h1 = __float2half_rn((float)v[2 * b2 + 1]);
                v[2 * b2] -= (double)__half2float(h0);
                v[2 * b2 + 1] -= (double)__half2float(h1);
                w[b2] = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
            }
            *reinterpret_cast<uint4 *>(Bs + pc * kOpPieceBytes + off) = make_uint4(w[0], w[1], w[2], w[3]);
        }
    }
    // K[8m + n] = L(D) = sum_l sum_a U[l][m][a] P[l][n][a]
    if(t < 64) {
        const int m = t >> 3, n = t & 7;
        double    k = 0.0;
        for(int l = 0; l < 8; l++)
#pragma unroll
            for(int a = 0; a < 8; a++) k += s_u[l][m][a] * s_p[l][n][a];
        // stored in the order the kernel's threads read it: a thread of class q4 = lane & 3 owns coefficients 8 j + 2 q4 + e
        // (j = 0..7, e = 0, 1) -- its 16 values lie together: position 16 q4 + 2 j + e
        p.op.K[(size_t)slot * 64 + 16 * ((t & 7) >> 1) + 2 * (t >> 3) + (t & 1)] = (float)k;
    }
    // the component's first slot also writes what depends on the tables only: 1/q (biased) and 512 q
    if(slot == p.drop.gtile_start[c] * 32 && t < 64) {
        p.op.rq[c * 64 + t] = quant_rcp_fast((float)s_q[t]);
        p.op.q512[c * 64 + t] = 512.0f * (float)s_q[t];
    }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------------------------------
// One CTA per SM, warp-specialised:
//   * kOpGroups groups of four COMPUTE warps.  Item = a PAIR of consecutive list slots (two dropon blocks, next to each
//     other in the image wherever the list entries are) x ALL images, items dealt round-robin to the CTAs; inside an item,
//     batches of 64 images, batch k to group k mod kOpGroups.  A unit = (item, batch) = 128 rows: rows 0..63 the first
//     slot's block of the 64 images, rows 64..127 the second slot's.  The pairing is what the memory system asks for: a
//     lone 128-byte block per image runs at 4.6 TB/s, two adjacent ones requested together at 5.9 TB/s
//     (profiles/microbench/gather_ubench.txt).  Each half is one UMMA chain of M = 64 against its slot's operator; the
//     second half's accumulator goes to lanes 16..31 of every 32-lane quarter, so warp w of the group holds both blocks of
//     images 16 w .. 16 w + 15 -- in shared memory, in tensor memory and in the stores.
//     A group's units form one stream across item boundaries, software-pipelined over three shared-memory stages and two
//     accumulators:  iteration u = convert unit u and hand it to the tensor core | requantise unit u - 1 (its UMMAs ran
//     meanwhile) and store it | request unit u + 2 into the stage that just became free.
//   * one MMA warp per group (one lane each): waits for the group's next unit, issues its UMMAs, commits to its mbarrier.
//   * one LOAD warp (one lane): brings the two operators of the next item into the other of two buffers (one bulk copy).
// No warp ever waits for a sibling: all hand-overs are mbarriers with the consumer one pipeline step behind the producer.
struct OpBars { // mbarrier indices
    static constexpr int kFull = 0;                                       // [2]  operators of an item have landed
    static constexpr int kEmpty = 2;                                      // [2]  every UMMA that read them has completed
    static constexpr int kConv = 4;                                       // [G][3] the unit's rows are converted (4 warps)
    static constexpr int kMma = kConv + kOpGroups * kOpStages;            // [G][2] the unit's UMMAs have completed
    static constexpr int kFree = kMma + kOpGroups * 2;                    // [G][2] the accumulator has been read (4 warps)
    static constexpr int kDrain = kFree + kOpGroups * 2;                  // [G]  end of the kernel
    static constexpr int kCount = kDrain + kOpGroups;
};

template <int NP>
struct OpSmem {
    static constexpr int kStage = 0;                                                   // [group][stage] 16 KB
    static constexpr int kB = kStage + kOpGroups * kOpStages * kOpStageBytes;          // [buf][slot of the pair][piece] 8 KB
    static constexpr int kBBuf = 2 * NP * kOpPieceBytes;
    static constexpr int kK = kB + 2 * kBBuf;                                          // [ring][slot of the pair][64] floats
    static constexpr int kRq = kK + kOpKRing * 512;                                    // [comp][64] biased 1/q
    static constexpr int kQ512 = kRq + MJX_MAX_COMPONENTS * 256;                       // [comp][64] 512 q
    static constexpr int kAddr = kQ512 + MJX_MAX_COMPONENTS * 256;                     // [warp][stage][32] global addresses
    static constexpr int kBar = kAddr + kOpGroups * 4 * kOpStages * 256;
    static constexpr int kMisc = kBar + 8 * OpBars::kCount;                            // tmem base, info[4]
    static constexpr int kBytes = kMisc + 64 + 1024;                                   // + alignment slack
    static_assert(kBytes <= 232448, "shared memory of one CTA");
};

__device__ __forceinline__ uint4 lds128(uint32_t a) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ float4 lds128f(uint32_t a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void sts128(uint32_t a, uint4 v) {
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ unsigned long long lds64(uint32_t a) {
    unsigned long long v;
    asm volatile("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void sts64(uint32_t a, unsigned long long v) { asm volatile("st.shared.u64 [%0], %1;" ::"r"(a), "l"(v) : "memory"); }
__device__ __forceinline__ void sts32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ F2 lds64f(uint32_t a) {
    float x, y;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(x), "=f"(y) : "r"(a) : "memory");
    return f2(x, y);
}
__device__ __forceinline__ uint32_t lds32(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}
// the two halves of a packed fp16 pair as floats
__device__ __forceinline__ F2 half2_to_f2(uint32_t h) {
    const float2 f = __half22float2(*reinterpret_cast<const __half2 *>(&h));
    return f2(f.x, f.y);
}

// one pair of coefficients: acc = -(S / 512) L(I o q) from the accumulator, i512 = I / 512 from the staged fp16 row, k = L(D):
//   y = k + acc * (512 / S)                  the reference's blend term Y (src/compose.c:300-312)
//   t = trunc(y), a = I*q + t, out = trunc(a / q) as int16 bits          (src/compose.c:315-336)
// Same devices as requant_pair (mjx_math.cuh): trunc(y) from one round-toward-zero add onto +-2^23, a assembled exactly
// from integers below 2^24, the division by one RZ FMA with the biased reciprocal.
__device__ __forceinline__ uint32_t requant_pair_op(F2 acc, float ascale, F2 k, F2 i512, F2 q512, F2 rq) {
    const F2 y = fma2(acc, bc2(ascale), k);
    const F2 sm = signed_magic2(y);
    const F2 u = add2_rz(y, sm);             // sm + trunc(y)
    const F2 v = fma2(i512, q512, neg2(sm)); // I*q - sm
    return tdiv_pair(add2(u, v), rq);
}

template <int NP, bool kCheck>
__global__ void __launch_bounds__(kOpThreads, 1) k2_generic_op_kernel(const OpParams p) {
    using L = OpSmem<NP>;
    constexpr int G = kOpGroups;
    extern __shared__ __align__(1024) unsigned char smem_dyn[];
    const uint32_t base32 = (smem_u32(smem_dyn) + 1023u) & ~1023u; // SWIZZLE_128B atoms are 1024 bytes
    const int      lane = threadIdx.x & 31, widx = threadIdx.x >> 5;
    const uint32_t sMisc = base32 + L::kMisc; // [0] tmem base, [4..7] info
    auto           bar = [&](int i) { return base32 + L::kBar + 8u * (uint32_t)i; };
    const int      n_pairs = p.drop.n_generic >> 1; // the list is padded to multiples of 32 slots per component: pairs never mix components
    const int      nb = (p.n + 63) >> 6;            // batches of 64 images

    // ---- one-time set-up ----
    if(widx == 0) tmem_alloc<512>(sMisc);
    if(threadIdx.x == 32) {
        const int active = nb < G ? nb : G; // groups that have units
        for(int i = 0; i < OpBars::kCount; i++) {
            const bool four = (i >= OpBars::kConv && i < OpBars::kMma) || (i >= OpBars::kFree && i < OpBars::kDrain);
            const bool per_group = i >= OpBars::kEmpty && i < OpBars::kConv;
            mbar_init(bar(i), four ? 4u : (per_group ? (uint32_t)active : 1u));
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if(threadIdx.x >= 64 && threadIdx.x < 64 + MJX_MAX_COMPONENTS) {
        const int c = threadIdx.x - 64;
        sts32(sMisc + 16 + 4 * c, c < p.drop.ncomp ? (uint32_t)p.op.info[c] : 0xffffffffu);
    }
    for(int i = threadIdx.x; i < MJX_MAX_COMPONENTS * 64; i += kOpThreads) {
        const float rq = i < p.drop.ncomp * 64 ? p.op.rq[i] : 1.0f, q512 = i < p.drop.ncomp * 64 ? p.op.q512[i] : 512.0f;
        asm volatile("st.shared.f32 [%0], %1;" ::"r"(base32 + L::kRq + 4 * i), "f"(rq) : "memory");
        asm volatile("st.shared.f32 [%0], %1;" ::"r"(base32 + L::kQ512 + 4 * i), "f"(q512) : "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = lds32(sMisc);
    auto           served = [&](uint32_t e) { return e != 0xffffffffu && (int)lds32(sMisc + 16 + 4 * entry_comp(e)) >= 0; };
    // the two list entries of this CTA's item j (static round-robin over the pairs); 0xffffffff: padding slot / past the end
    auto peek = [&](int j) -> uint2 {
        const int pr = blockIdx.x + j * gridDim.x;
        if(pr >= n_pairs) return make_uint2(0xffffffffu, 0xffffffffu);
        return __ldg(reinterpret_cast<const uint2 *>(p.drop.list_generic) + pr);
    };
    // this CTA's next item after item j with a block the kernel serves; returns its entries, both 0xffffffff at the end
    auto next_item = [&](int &j) -> uint2 {
        for(;;) {
            j++;
            if(blockIdx.x + j * gridDim.x >= n_pairs) return make_uint2(0xffffffffu, 0xffffffffu);
            uint2 e = peek(j);
            if(!served(e.x)) e.x = 0xffffffffu;
            if(!served(e.y)) e.y = 0xffffffffu;
            if((e.x & e.y) != 0xffffffffu) return e;
        }
    };

    if(widx >= G * 4 && widx < G * 5) {
        // ================= MMA warps: one per group =================
        // One lane per group waits for the group's next unit (its four warps have converted it, the accumulator it goes to has
        // been read, the item's operators have landed), issues the unit's UMMAs and commits them to the unit's mbarrier.
        const int g = widx - G * 4;
        const int bpg = g < nb ? (nb - g + G - 1) / G : 0;
        if(lane == 0 && bpg > 0) {
            int j = -1, jv = 0, u = 0;
            for(uint2 e = next_item(j); (e.x & e.y) != 0xffffffffu; e = next_item(j), jv++) {
                const int buf = jv & 1;
                mbar_wait(bar(OpBars::kFull + buf), (jv >> 1) & 1);
                const uint64_t bd0 = umma_desc_sw128(base32 + L::kB + buf * L::kBBuf);
                for(int k = 0; k < bpg; k++, u++) {
                    const int st = u % kOpStages, acc = u & 1;
                    // (busy polling: a suspended wait wakes this lane too late -- 1.51 instead of 1.46 ms on the bench batch)
                    mbar_spin(bar(OpBars::kConv + g * kOpStages + st), (u / kOpStages) & 1);
                    if(u >= 2) mbar_spin(bar(OpBars::kFree + g * 2 + acc), ((u >> 1) - 1) & 1);
                    tc_fence_after();
                    const uint64_t ad0 = umma_desc_sw128(base32 + L::kStage + (g * kOpStages + st) * kOpStageBytes);
                    const uint32_t tmem_d0 = tmem_base + g * 128 + acc * 64;
#pragma unroll
                    for(int half = 0; half < 2; half++) {
                        // rows 64 half .. 64 half + 63 against the operator of the pair's slot `half`; M = 64 puts row i on lane
                        // 32 (i / 16) + i % 16 of the accumulator, the second half starts at lane 16
                        const uint32_t tmem_d = tmem_d0 + ((uint32_t)(16 * half) << 16);
#pragma unroll
                        for(int pc = 0; pc < NP; pc++)
#pragma unroll
                            for(int ks = 0; ks < 4; ks++)
                                umma_f16(tmem_d, ad0 + (uint64_t)(half * (8192 >> 4) + 2 * ks), bd0 + (uint64_t)((half * NP + pc) * (kOpPieceBytes >> 4) + 2 * ks),
                                         umma_idesc_f16(64, 64), (pc | ks) != 0);
                    }
                    umma_commit(bar(OpBars::kMma + g * 2 + acc));
                }
                umma_commit(bar(OpBars::kEmpty + buf)); // this group's UMMAs on the item's operators; the barrier counts the groups
            }
            umma_commit(bar(OpBars::kDrain + g));
            mbar_wait(bar(OpBars::kDrain + g), 0); // no commit of this lane is left in flight when the CTA ends
        }
    }
    else if(widx == G * 5) {
        // ================= LOAD warp =================
        if(lane == 0) {
            int j = -1, jv = 0;
            for(uint2 e = next_item(j); (e.x & e.y) != 0xffffffffu; e = next_item(j), jv++) {
                const int buf = jv & 1;
                if(jv >= 2) mbar_wait(bar(OpBars::kEmpty + buf), ((jv >> 1) - 1) & 1);
                const size_t   slot0 = 2 * (size_t)(blockIdx.x + j * gridDim.x);
                const uint32_t full = bar(OpBars::kFull + buf);
                mbar_arrive_expect_tx(full, 2 * NP * kOpPieceBytes + 512);
                // the two slots' pieces lie back to back in the cache: one copy (a padding slot's part is never looked at)
                bulk_g2s(base32 + L::kB + buf * L::kBBuf, p.op.B + slot0 * NP * kOpPieceBytes, 2 * NP * kOpPieceBytes, full);
                bulk_g2s(base32 + L::kK + (jv & (kOpKRing - 1)) * 512, p.op.K + slot0 * 64, 512, full);
            }
        }
    }
    else {
        // ================= compute warps =================
        const int grp = widx >> 2, wq = widx & 3;
        const int bpg = grp < nb ? (nb - grp + G - 1) / G : 0; // batches of this group per item
        if(bpg > 0) {
            const uint32_t taddr = tmem_base + grp * 128 + ((uint32_t)(wq * 32) << 16); // this warp's lanes of the group's two accumulators
            // This warp's 32 rows of a unit: images 16 wq .. 16 wq + 15 of the batch, first slot (tile rows 16 wq + i) and second
            // slot (tile rows 64 + 16 wq + i).  Tile row r lies at r * 128, chunk c of it at (c ^ (r & 7)) * 16.
            const int      half = lane >> 4, idx = lane & 15; // thread-per-row view (addresses, conversion): my row
            const uint32_t stage0 = base32 + L::kStage + (grp * kOpStages) * kOpStageBytes;
            const uint32_t my_row0 = stage0 + half * 8192 + (16 * wq + idx) * 128;
            const int      rsw = lane & 7;
            // 8-lanes-per-row view (cp.async in, coalesced stores out): instruction i moves rows (image 2 i + (lane >> 4), slot
            // (lane >> 3) & 1): both blocks of two images -- 2 x 256 contiguous bytes where the slots are neighbours
            const int      g_half = (lane >> 3) & 1, g_img = lane >> 4;
            const uint32_t g_row0 = stage0 + g_half * 8192 + (16 * wq + g_img) * 128; // + i * 256; chunk (lane & 7) ^ ((2 i + g_img) & 7)
            const uint32_t g_addr0 = (uint32_t)(16 * g_half + g_img) * 8;             // + i * 16: where that row's global address sits
            const uint32_t waddr = base32 + L::kAddr + widx * kOpStages * 256;
            uint32_t       kx; // (w & 0x07FF07FF) ^ kx as ONE LOP3: the constant must live in a register
            asm volatile("mov.u32 %0, 0x04000400;" : "=r"(kx));

            // a position in the group's stream of units: item j (the jv-th served one, list entries e), batch grp + k * G;
            // en = list entries of item j + 1, requested a whole item before they are looked at
            struct Pos {
                int   j, jv, k;
                uint2 e, en;
            };
            auto at_end = [](const Pos &q) { return (q.e.x & q.e.y) == 0xffffffffu; };
            auto advance = [&](Pos &q) {
                if(at_end(q)) return;
                if(++q.k < bpg) return;
                q.k = 0;
                q.jv++;
                q.j++;
                q.e = q.en;
                if(!served(q.e.x)) q.e.x = 0xffffffffu;
                if(!served(q.e.y)) q.e.y = 0xffffffffu;
                if(at_end(q)) q.e = next_item(q.j); // padding pair, component not served, or the end: look further (rare)
                q.en = peek(q.j + 1);
            };
            auto comp_of = [](const Pos &q) { return entry_comp(q.e.x != 0xffffffffu ? q.e.x : q.e.y); };
            auto image_of = [&](const Pos &q) { return (grp + q.k * G) * 64 + 16 * wq + idx; };
            auto table_of = [&](const Pos &q) -> uint4 {
                const int img = image_of(q);
                if(at_end(q) || img >= p.n) return make_uint4(0u, 0u, 0u, 0u);
                return __ldg(p.table + (size_t)comp_of(q) * p.n + img);
            };
            // request the warp's 32 rows of the unit at q into stage st; t = this thread's table entry of that unit
            auto prefetch = [&](const Pos &q, const uint4 t, int st) {
                unsigned long long a = 0ull;
                const uint32_t     e = half ? q.e.y : q.e.x;
                if(e != 0xffffffffu) {
                    const DropComp &dc = p.drop.comp[entry_comp(e)];
                    const uint32_t  row = (uint32_t)(p.block_y * dc.vs + entry_row(e)), col = (uint32_t)(p.block_x * dc.hs + entry_col(e));
                    const unsigned long long plane = (unsigned long long)t.x | ((unsigned long long)t.y << 32);
                    if(plane != 0ull && row < t.w && col < t.z) a = plane + ((unsigned long long)row * t.z + col) * 128ull;
                }
                const uint32_t ad = waddr + st * 256;
                sts64(ad + lane * 8, a);
                __syncwarp();
                if(at_end(q)) return; // warp-uniform: past the end of the stream
                const uint32_t     dst = g_row0 + st * kOpStageBytes;
                const unsigned     coff = (lane & 7) * 16;
                unsigned long long b[8];
#pragma unroll
                for(int i = 0; i < 8; i++) b[i] = lds64(ad + g_addr0 + i * 16);
#pragma unroll
                for(int i = 0; i < 8; i++)
                    cp_async16_cg(dst + i * 256 + (((lane & 7) ^ ((2 * i + g_img) & 7)) << 4), reinterpret_cast<const void *>((b[i] ? b[i] : (unsigned long long)(uintptr_t)p.items) + coff),
                               b[i] ? 16u : 0u);
            };

            Pos pc, pp; // unit being converted, unit being requested (two ahead)
            pc.j = -1, pc.jv = 0, pc.k = 0;
            pc.e = next_item(pc.j);
            pc.en = peek(pc.j + 1);
            pp = pc;
            // prologue: units 0, 1, 2 into stages 0, 1, 2
#pragma unroll
            for(int i = 0; i < kOpStages; i++) {
                prefetch(pp, table_of(pp), i);
                cp_async_commit();
                advance(pp);
            }
            Pos pe = pc; // unit being requantised (one behind pc)
            // what the requantisation needs per coefficient, for this thread's 8 pairs (columns 8 j + 2 q4 + {0, 1}): 1/q and
            // 512 q of the component stay in registers, K of the unit's two slots is fetched per unit (8 bytes per pair)
            const int q4 = lane & 3, r8 = lane >> 2;
            F2        rq2[8], q5122[8];
            float     ascale = 1.0f;
            int       c_regs = -1;
#pragma unroll
            for(int j = 0; j < 8; j++) rq2[j] = q5122[j] = f2(0.f, 0.f);
            for(int u = 0;; u++) {
                const bool conv = !at_end(pc); // unit u exists
                if(!conv && u == 0) break;
                const uint4 tnext = table_of(pp); // for the request at the end of this iteration
                if(conv) {
                    const int st = u % kOpStages;
                    if(u == 0) cp_async_wait<2>();
                    else cp_async_wait<1>(); // everything but the newest request: this unit's rows have landed
                    __syncwarp();
                    // ---- int16 -> fp16 (I / 512) in place, thread per row ----
                    const uint32_t my_row = my_row0 + st * kOpStageBytes;
                    uint32_t       vmax = 0x80008000u, vmin = 0x7fff7fffu; // running max / min of the row's coefficients, two int16 lanes
                    uint4          w[8];
#pragma unroll
                    for(int ch = 0; ch < 8; ch++) w[ch] = lds128(my_row + ((ch ^ rsw) << 4));
#pragma unroll
                    for(int ch = 0; ch < 8; ch++) {
                        uint32_t *pw = &w[ch].x;
                        if(kCheck) { // one packed three-input max and one min per two words (VIMNMX3.S16x2)
                            vmax = __vimax3_s16x2(vmax, pw[0], pw[1]), vmax = __vimax3_s16x2(vmax, pw[2], pw[3]);
                            vmin = __vimin3_s16x2(vmin, pw[0], pw[1]), vmin = __vimin3_s16x2(vmin, pw[2], pw[3]);
                        }
#pragma unroll
                        for(int i = 0; i < 4; i++) {
                            uint32_t x;
                            asm("lop3.b32 %0, %1, 0x07FF07FF, %2, 0x6A;" : "=r"(x) : "r"(pw[i]), "r"(kx)); // (w & mask) ^ kx
                            asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(pw[i]) : "r"(x), "r"(0x78007800u), "r"(0xC000C000u));
                        }
                        sts128(my_row + ((ch ^ rsw) << 4), w[ch]);
                    }
                    if(kCheck) {
                        const uint32_t ad = waddr + st * 256 + lane * 8;
                        const int      hi = max((int)(int16_t)(vmax & 0xffffu), (int)vmax >> 16), lo = min((int)(int16_t)(vmin & 0xffffu), (int)vmin >> 16);
                        if((hi > 1023 || lo < -1024) && lds64(ad) != 0ull) { // leave this block to the fp32 kernel: one bit per (list slot, image), and do not store it
                            const int slot = 2 * (blockIdx.x + pc.j * gridDim.x) + half;
                            atomicOr(p.redo_mask + (size_t)(slot >> 5) * p.n + image_of(pc), 1u << (slot & 31));
                            atomicAdd(p.redo_count, 1u);
                            sts64(ad, 0ull);
                        }
                    }
                    fence_async_smem(); // the rows, as the tensor core will read them
                    __syncwarp();
                    if(lane == 0) mbar_arrive(bar(OpBars::kConv + grp * kOpStages + st));
                }
                if(u > 0) {
                    // ---- requantise unit u - 1: its UMMAs ran while the previous unit was stored and this one converted ----
                    // The accumulator is read in the fragment layout (16x256b): this thread sees the SAME 16 coefficients
                    // (8 pairs) of four rows, so what depends on the coefficient only -- K of the item's slots, 1/q and 512 q of
                    // the component -- lives in registers and is reloaded when the item / the component changes, not per unit.
                    const int      v = u - 1, st = v % kOpStages, acc = v & 1, c = comp_of(pe);
                    const uint32_t ad = waddr + st * 256;
                    if(c != c_regs) {
                        c_regs = c;
                        ascale = (float)(1 << (9 - (int)lds32(sMisc + 16 + 4 * c)));
#pragma unroll
                        for(int j = 0; j < 8; j++) {
                            rq2[j] = lds64f(base32 + L::kRq + c * 256 + (8 * j + 2 * q4) * 4);
                            q5122[j] = lds64f(base32 + L::kQ512 + c * 256 + (8 * j + 2 * q4) * 4);
                        }
                    }
                    mbar_wait(bar(OpBars::kMma + grp * 2 + acc), (v >> 1) & 1); // (also: the item's K have landed)
                    tc_fence_after();
                    const uint32_t sK = base32 + L::kK + (pe.jv & (kOpKRing - 1)) * 512 + q4 * 64;
                    // word (q4) of chunk j of tile row 64 h + 16 wq + r8 + 8 s2: the fp16 pair I / 512 on the way in, the int16 pair on
                    // the way out; lanes 16 h .. 16 h + 15 of the quarter hold half h
                    const uint32_t my_w = stage0 + st * kOpStageBytes + (16 * wq + r8) * 128 + q4 * 4;
#pragma unroll
                    for(int h = 0; h < 2; h++) {
                        float y[32];
                        tmem_ld_16x256b_x8(taddr + ((uint32_t)(16 * h) << 16) + acc * 64, y);
                        F2 k2[8]; // K of the slot this half belongs to (stored per thread class: 64 contiguous bytes)
#pragma unroll
                        for(int j4 = 0; j4 < 4; j4++) {
                            const float4 kk = lds128f(sK + h * 256 + j4 * 16);
                            k2[2 * j4] = f2(kk.x, kk.y), k2[2 * j4 + 1] = f2(kk.z, kk.w);
                        }
                        tmem_wait_ld();
                        if(h == 1) {
                            tc_fence_before();
                            __syncwarp();
                            if(lane == 0) mbar_arrive(bar(OpBars::kFree + grp * 2 + acc)); // the accumulator may be overwritten
                        }
                        // all loads, then all arithmetic (16 independent chains the compiler can interleave), then all stores
                        uint32_t io[16];
#pragma unroll
                        for(int j = 0; j < 8; j++)
#pragma unroll
                            for(int s2 = 0; s2 < 2; s2++) io[2 * j + s2] = lds32(my_w + h * 8192 + s2 * 1024 + ((j ^ r8) << 4));
#pragma unroll
                        for(int j = 0; j < 8; j++)
#pragma unroll
                            for(int s2 = 0; s2 < 2; s2++)
                                io[2 * j + s2] = requant_pair_op(f2(y[4 * j + 2 * s2], y[4 * j + 2 * s2 + 1]), ascale, k2[j], half2_to_f2(io[2 * j + s2]), q5122[j], rq2[j]);
#pragma unroll
                        for(int j = 0; j < 8; j++)
#pragma unroll
                            for(int s2 = 0; s2 < 2; s2++) sts32(my_w + h * 8192 + s2 * 1024 + ((j ^ r8) << 4), io[2 * j + s2]);
                    }
                    __syncwarp();
                    // ---- coalesced write-back, the gather's mapping in reverse: both blocks of two images per instruction ----
                    {
                        const uint32_t     src = g_row0 + st * kOpStageBytes;
                        const unsigned     coff = (lane & 7) * 16;
                        unsigned long long b[8];
                        uint4              w[8];
#pragma unroll
                        for(int i = 0; i < 8; i++) {
                            b[i] = lds64(ad + g_addr0 + i * 16);
                            w[i] = lds128(src + i * 256 + (((lane & 7) ^ ((2 * i + g_img) & 7)) << 4));
                        }
#pragma unroll
                        for(int i = 0; i < 8; i++)
                            if(b[i]) __stcs(reinterpret_cast<uint4 *>(b[i] + coff), w[i]);
                    }
                    __syncwarp();
                    // ---- request unit u + 2 into the stage that just became free ----
                    prefetch(pp, tnext, st);
                    cp_async_commit();
                    advance(pp);
                    advance(pe);
                }
                if(!conv) break; // the last unit has been requantised
                advance(pc);
            }
            cp_async_wait<0>();
        }
    }
    // every warp is done with its tensor-memory lanes before warp 0 returns the allocation
    tc_fence_before();
    __syncthreads();
    if(widx == 0) tmem_dealloc<512>(tmem_base);
}

// ---------------------------------------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------------------------------------

size_t op_cache_bytes(int n_generic, int np, OpView *v) {
    size_t off = 0;
    auto   take = [&](size_t bytes) {
        const size_t o = off;
        off = (off + bytes + 1023) / 1024 * 1024;
        return o;
    };
    const size_t oB = take((size_t)n_generic * np * kOpPieceBytes);
    const size_t oK = take((size_t)n_generic * 256);
    const size_t oD = take(MJX_MAX_COMPONENTS * 256);
    const size_t oR = take(MJX_MAX_COMPONENTS * 256);
    const size_t oKey = take(MJX_MAX_COMPONENTS * 128);
    const size_t oI = take(MJX_MAX_COMPONENTS * 4);
    const size_t oRb = take(MJX_MAX_COMPONENTS * 4);
    if(v) {
        unsigned char *b = v->B; // the caller put the slab's base here
        v->B = b + oB;
        v->K = reinterpret_cast<float *>(b + oK);
        v->q512 = reinterpret_cast<float *>(b + oD);
        v->rq = reinterpret_cast<float *>(b + oR);
        v->key = reinterpret_cast<uint16_t *>(b + oKey);
        v->info = reinterpret_cast<int *>(b + oI);
        v->rebuild = reinterpret_cast<int *>(b + oRb);
        v->np = np;
    }
    return off;
}

template <int NP>
static cudaError_t launch_op_kernel(cudaStream_t s, const OpParams &p, int sms, bool check, bool *attr_set) {
    using L = OpSmem<NP>;
    cudaError_t e;
    if(!*attr_set) {
        if((e = cudaFuncSetAttribute(k2_generic_op_kernel<NP, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kBytes)) != cudaSuccess) return e;
        if((e = cudaFuncSetAttribute(k2_generic_op_kernel<NP, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kBytes)) != cudaSuccess) return e;
        *attr_set = true;
    }
    const int pairs = p.drop.n_generic / 2;
    const int ctas = pairs < sms ? pairs : sms;
    if(check) k2_generic_op_kernel<NP, true><<<ctas, kOpThreads, L::kBytes, s>>>(p);
    else k2_generic_op_kernel<NP, false><<<ctas, kOpThreads, L::kBytes, s>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_k2_generic_op(cudaStream_t s, const OpParams &p, int sm_count, bool check, bool *attr_set, int *launches) {
    cudaError_t e;
    const int   sms = sm_count > 0 ? sm_count : 148;
    if(p.op.np != 2) return cudaErrorInvalidValue;
    k2_op_prepare_kernel<<<(p.n * p.drop.ncomp + 127) / 128, 128, 0, s>>>(p);
    if((e = cudaGetLastError()) != cudaSuccess) return e;
    // one CTA per SM walks the slots; when the tables are unchanged every CTA leaves at once
    const int bctas = p.drop.n_generic < sms ? p.drop.n_generic : sms;
    k2_op_build_kernel<2><<<bctas, 256, 0, s>>>(p);
    if((e = cudaGetLastError()) != cudaSuccess) return e;
    if((e = launch_op_kernel<2>(s, p, sms, check, attr_set)) != cudaSuccess) return e;
    if(launches) *launches += 3;
    return cudaSuccess;
}

} // namespace mjx
