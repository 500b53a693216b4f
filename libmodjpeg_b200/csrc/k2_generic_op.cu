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
static constexpr int kOpStages = 2;
static constexpr int kOpPieceBytes = 64 * 128;  // 64 output rows x 64 fp16 (K-major)
static constexpr int kOpKRing = 4;

template <int NP, int G>
struct OpSmem {
    static constexpr int kStage = 0;                                             // [group][stage] 16 KB
    static constexpr int kB = kStage + G * kOpStages * kOpStageBytes;            // [buf]: hi tile (T part + diag part), lo pieces
    static constexpr int kBBuf = 2 * kOpPieceBytes + (NP - 1) * kOpPieceBytes;
    static constexpr int kK = kB + 2 * kBBuf;                                    // [ring][64] floats
    static constexpr int kRq = kK + kOpKRing * 256;                              // [comp][64] floats
    static constexpr int kAddr = kRq + MJX_MAX_COMPONENTS * 256;                 // [warp][stage][32] global addresses
    static constexpr int kBar = kAddr + G * 4 * kOpStages * 256;                 // mma[G], full[2]
    static constexpr int kMisc = kBar + 8 * (G + 2);                             // tmem base, done[2], info[4]
    static constexpr int kBytes = kMisc + 64 + 1024;                             // + alignment slack
};

// one pair of coefficients from the accumulator: acc = -(S / 512) L(I o q), iqs = I*q / 512 (both exact products of the
// tensor core), k = L(D):
//   y = k + acc * (512 / S)                  the reference's blend term Y (src/compose.c:300-312)
//   t = trunc(y), a = I*q + t, out = trunc(a / q) as int16 bits          (src/compose.c:315-336)
// Same devices as requant_pair (mjx_math.cuh): trunc(y) from one round-toward-zero add onto +-2^23, a assembled exactly
// from integers below 2^24, the division by one RZ FMA with the biased reciprocal.
__device__ __forceinline__ uint32_t requant_pair_acc(F2 acc, float ascale, F2 k, F2 iqs, F2 rq) {
    const F2 y = fma2(acc, bc2(ascale), k);
    const F2 sm = signed_magic2(y);
    const F2 u = add2_rz(y, sm);                       // sm + trunc(y)
    const F2 v = fma2(iqs, bc2(512.0f), neg2(sm));     // I*q - sm
    return tdiv_pair(add2(u, v), rq);
}

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
    if(threadIdx.x < 32) {
        for(int c = 0; c < ncomp; c++) {
            const uint32_t w = reinterpret_cast<const uint32_t *>(&p.items[0].q[c][0])[threadIdx.x];
            s_q[c][threadIdx.x] = w;
            unsigned qm = max(w & 0xffffu, w >> 16), qn = min(w & 0xffffu, w >> 16);
            qm = __reduce_max_sync(0xffffffffu, qm);
            qn = __reduce_min_sync(0xffffffffu, qn);
            // scale S = 2^sh: the largest power of two <= 512 that keeps |S q T| inside fp16 (|T| <= max alpha < 1.01)
            int sh = 9;
            while(sh > 0 && 1.01f * (float)qm * (float)(1 << sh) > 65000.0f) sh--;
            const int ok = (qm <= 255u && qn >= 1u) ? sh : -1;
            if(threadIdx.x == 0) s_ok[c] = ok;
            if(blockIdx.x == 0) {
                const uint32_t old = reinterpret_cast<const uint32_t *>(p.op.key + c * 64)[threadIdx.x];
                const bool     differ = __any_sync(0xffffffffu, old != w);
                reinterpret_cast<uint32_t *>(p.op.key + c * 64)[threadIdx.x] = w;
                if(threadIdx.x == 0) {
                    p.op.rebuild[c] = differ ? 1 : 0;
                    p.op.info[c] = ok;
                }
            }
        }
    }
    __syncthreads();
    const int i = blockIdx.x * 128 + threadIdx.x;
    if(i >= p.n) return;
    const mjx_image_desc_t &im = p.items[i];
    const int               ntiles = p.drop.n_generic >> 5;
    for(int c = 0; c < ncomp; c++) {
        const int t0 = p.drop.gtile_start[c], t1 = c + 1 < ncomp ? p.drop.gtile_start[c + 1] : ntiles;
        bool      same = s_ok[c] >= 0;
        if(same && i > 0) {
            const uint4 *q = reinterpret_cast<const uint4 *>(&im.q[c][0]);
#pragma unroll
            for(int k = 0; k < 8; k++) {
                const uint4 a = __ldg(q + k);
                same = same && a.x == s_q[c][4 * k] && a.y == s_q[c][4 * k + 1] && a.z == s_q[c][4 * k + 2] && a.w == s_q[c][4 * k + 3];
            }
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
    const int      slot = blockIdx.x;
    const uint32_t e = __ldg(p.drop.list_generic + slot);
    if(e == 0xffffffffu) return;
    const int c = entry_comp(e);
    if(!p.op.rebuild[c]) return;
    const int sh = p.op.info[c];
    if(sh < 0) return;
    const DropComp &dc = p.drop.comp[c];
    const size_t    bi = (size_t)entry_row(e) * dc.wb + entry_col(e);
    const int       t = threadIdx.x;
    for(int x = t; x < 512; x += 256) s_m[x >> 6][(x >> 3) & 7][x & 7] = conv_m(x >> 6, (x >> 3) & 7, x & 7);
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
        p.op.K[(size_t)slot * 64 + t] = (float)k;
    }
    // the component's first slot also writes what depends on the tables only: the diag(q) half of the hi tile and 1/q
    if(slot == p.drop.gtile_start[c] * 32) {
        unsigned char *dg = p.op.diag + c * kOpPieceBytes;
        for(int x = t; x < 512; x += 256) {
            const int row = x >> 3, ch = x & 7; // row = coefficient, ch = 16-byte chunk of the row holding k = 8 ch .. 8 ch + 7
            uint32_t  w[4] = {0u, 0u, 0u, 0u};
            if(ch == (row >> 3)) {
                const uint32_t h = (uint32_t)__half_as_ushort(__float2half_rn((float)s_q[row])); // q <= 255: exact
                w[(row & 7) >> 1] = (row & 1) ? h << 16 : h;
            }
            *reinterpret_cast<uint4 *>(dg + sw128_off(row, 8 * ch)) = make_uint4(w[0], w[1], w[2], w[3]);
        }
        if(t < 64) p.op.rq[c * 64 + t] = quant_rcp_fast((float)s_q[t]);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------------------------------
// One CTA per SM, G groups of four warps.  Item = one list slot (dropon block) x ALL images, items dealt round-robin to
// the CTAs; inside an item, batches of 128 images (one UMMA each), batch k to group k mod G.  A thread owns one image of
// its group's batch: it gathers that image's block, converts it, and after the group's UMMA requantises and stores it.
// A group's batches form one stream across item boundaries (the next item's first batch is prefetched during this
// item's last one); only the operator pieces are per item: two shared-memory buffers, filled by the bulk-copy engine,
// handed over by the last group that finishes with the item two back.
template <int NP, int G, bool kCheck>
__global__ void __launch_bounds__(G * 128, 1) k2_generic_op_kernel(const OpParams p) {
    using L = OpSmem<NP, G>;
    extern __shared__ unsigned char smem_dyn[];
    unsigned char *base = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~static_cast<uintptr_t>(1023));
    const int      lane = threadIdx.x & 31, widx = threadIdx.x >> 5, grp = widx >> 2, wq = widx & 3;
    const uint32_t base32 = smem_u32(base);
    uint32_t      *sMisc = reinterpret_cast<uint32_t *>(base + L::kMisc); // [0] tmem base, [1..2] done counters, [4..7] info
    const float   *sRq = reinterpret_cast<const float *>(base + L::kRq);
    const uint32_t bar_mma = base32 + L::kBar + 8 * grp;
    const uint32_t bar_full0 = base32 + L::kBar + 8 * G;
    const int      n_slots = p.drop.n_generic;
    const int      nb = (p.n + 127) >> 7;
    const int      bpg = grp < nb ? (nb - grp + G - 1) / G : 0; // batches of this group per item

    // issue the operator pieces of this CTA's item jj into buffer jj & 1 (one thread)
    auto issue_item = [&](int jj) {
        const int slot = blockIdx.x + jj * gridDim.x;
        if(slot >= n_slots) return;
        const uint32_t bar = bar_full0 + 8 * (jj & 1);
        const uint32_t e = __ldg(p.drop.list_generic + slot);
        const int      c = entry_comp(e);
        if(e == 0xffffffffu || (int)sMisc[4 + c] < 0) { // nothing to load: the item is skipped by every group
            mbar_arrive(bar);
            return;
        }
        const uint32_t dst = base32 + L::kB + (jj & 1) * L::kBBuf;
        mbar_arrive_expect_tx(bar, (NP + 1) * kOpPieceBytes + 256);
        const unsigned char *src = p.op.B + (size_t)slot * NP * kOpPieceBytes;
        bulk_g2s(dst, src, kOpPieceBytes, bar);
        bulk_g2s(dst + kOpPieceBytes, p.op.diag + c * kOpPieceBytes, kOpPieceBytes, bar);
#pragma unroll
        for(int pc = 1; pc < NP; pc++) bulk_g2s(dst + (pc + 1) * kOpPieceBytes, src + pc * kOpPieceBytes, kOpPieceBytes, bar);
        bulk_g2s(base32 + L::kK + (jj & (kOpKRing - 1)) * 256, p.op.K + (size_t)slot * 64, 256, bar);
    };

    // ---- one-time set-up ----
    if(widx == 0) tmem_alloc<512>(smem_u32(&sMisc[0]));
    if(threadIdx.x == 32) {
#pragma unroll
        for(int g = 0; g < G + 2; g++) mbar_init(base32 + L::kBar + 8 * g, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        sMisc[1] = 0, sMisc[2] = 0;
    }
    if(threadIdx.x >= 64 && threadIdx.x < 64 + MJX_MAX_COMPONENTS) sMisc[4 + threadIdx.x - 64] = threadIdx.x - 64 < p.drop.ncomp ? (uint32_t)p.op.info[threadIdx.x - 64] : 0xffffffffu;
    for(int i = threadIdx.x; i < MJX_MAX_COMPONENTS * 64; i += G * 128) reinterpret_cast<float *>(base + L::kRq)[i] = i < p.drop.ncomp * 64 ? p.op.rq[i] : 1.0f;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if(threadIdx.x == 0) {
        issue_item(0);
        issue_item(1);
    }
    const uint32_t tmem_g = sMisc[0] + grp * 128;                // this group's accumulator: 64 blend columns + 64 product columns
    const uint32_t taddr = tmem_g + ((uint32_t)(wq * 32) << 16); // this warp's lanes of it
    uint32_t       mma_phase = 0;

    // the last group to finish with item j's operator pieces refills the buffer with item j + 2's
    auto handoff = [&](int j) {
        if(wq == 0 && lane == 0) {
            const uint32_t old = atomicAdd(&sMisc[1 + (j & 1)], 1u);
            if(old == (uint32_t)(G - 1)) {
                sMisc[1 + (j & 1)] = 0;
                issue_item(j + 2);
            }
        }
    };

    if(bpg == 0) { // more groups than batches: one warp of this group takes part in the buffer protocol, nothing else
        if(wq == 0)
            for(int j = 0; blockIdx.x + j * gridDim.x < n_slots; j++) {
                mbar_wait(bar_full0 + 8 * (j & 1), (j >> 1) & 1);
                handoff(j);
            }
    }
    else {
        // staging layout.  Row r of the warp's 32 lies at r * 128, chunk c of row r at (c ^ (r & 7)) * 16.
        unsigned char *wstage0 = base + L::kStage + (grp * kOpStages) * kOpStageBytes + wq * 4096; // stage 0 of this warp
        const uint32_t wstage0_32 = smem_u32(wstage0);
        const int      rsw = lane & 7; // thread-per-row view (conversion, results): my row = lane
        // 8-lanes-per-row view (cp.async in, coalesced stores out): lane -> chunk (lane & 7) of rows (lane >> 3) + 4 j
        const uint32_t cp_off0 = (uint32_t)(lane >> 3) * 128 + (uint32_t)(((lane & 7) ^ (lane >> 3)) << 4);       // j even
        const uint32_t cp_off1 = (uint32_t)(lane >> 3) * 128 + (uint32_t)(((lane & 7) ^ ((lane >> 3) + 4)) << 4); // j odd
        unsigned long long *waddr = reinterpret_cast<unsigned long long *>(base + L::kAddr) + widx * kOpStages * 32;
        uint32_t kx; // (w & 0x07FF07FF) ^ kx as ONE LOP3: the constant must live in a register
        asm volatile("mov.u32 %0, 0x04000400;" : "=r"(kx));

        // unit u of this group = (item u / bpg, batch grp + (u % bpg) * G).  What a unit needs before its rows can be requested:
        struct Unit {
            uint32_t e;   // list entry of the item (0xffffffff: none / skipped)
            int      img; // this thread's image
        };
        auto unit_of = [&](int u) {
            Unit      un;
            const int j = u / bpg, k = u - j * bpg;
            const int slot = blockIdx.x + j * gridDim.x;
            un.e = slot < n_slots ? __ldg(p.drop.list_generic + slot) : 0xffffffffu;
            if(un.e != 0xffffffffu && (int)sMisc[4 + entry_comp(un.e)] < 0) un.e = 0xffffffffu;
            un.img = (grp + k * G) * 128 + wq * 32 + lane;
            return un;
        };
        // address of this thread's block of the unit (0: absent), from the compact table
        auto block_addr = [&](const Unit &un) -> unsigned long long {
            if(un.e == 0xffffffffu || un.img >= p.n) return 0ull;
            const int       c = entry_comp(un.e);
            const DropComp &dc = p.drop.comp[c];
            const uint4     t = __ldg(p.table + (size_t)c * p.n + un.img);
            const uint32_t  row = (uint32_t)(p.block_y * dc.vs + entry_row(un.e)), col = (uint32_t)(p.block_x * dc.hs + entry_col(un.e));
            const unsigned long long plane = (unsigned long long)t.x | ((unsigned long long)t.y << 32);
            if(plane == 0 || row >= t.w || col >= t.z) return 0ull;
            return plane + ((unsigned long long)row * t.z + col) * 128ull;
        };
        // request the warp's 32 rows of a unit into stage st
        auto prefetch = [&](const Unit &un, int st) {
            unsigned long long *addr = waddr + st * 32;
            const unsigned long long a = block_addr(un);
            addr[lane] = a;
            __syncwarp();
            if(un.e == 0xffffffffu) return; // warp-uniform
            const uint32_t            dst = wstage0_32 + st * kOpStageBytes;
            const unsigned long long *ap = addr + (lane >> 3);
            const unsigned            coff = (lane & 7) * 16;
            unsigned long long        b[8];
#pragma unroll
            for(int j = 0; j < 8; j++) b[j] = ap[4 * j];
#pragma unroll
            for(int j = 0; j < 8; j++)
                cp_async16(dst + j * 512 + ((j & 1) ? cp_off1 : cp_off0), reinterpret_cast<const void *>((b[j] ? b[j] : (unsigned long long)(uintptr_t)p.items) + coff),
                           b[j] ? 16u : 0u);
        };

        Unit cur = unit_of(0);
        prefetch(cur, 0);
        cp_async_commit();
        for(int u = 0;; u++) {
            const int j = u / bpg, k = u - j * bpg, st = u & 1;
            if(blockIdx.x + j * gridDim.x >= n_slots) break;
            const Unit nxt = unit_of(u + 1);
            prefetch(nxt, st ^ 1);
            cp_async_commit();
            if(k == 0) mbar_wait(bar_full0 + 8 * (j & 1), (j >> 1) & 1); // the item's operator pieces have landed
            cp_async_wait<1>();                                            // everything but the newest group: this unit's rows have landed
            __syncwarp();

            if(cur.e != 0xffffffffu) { // warp-uniform (in fact CTA-uniform)
                const int      c = entry_comp(cur.e);
                unsigned char *my_rowp = wstage0 + st * kOpStageBytes + lane * 128;
                uint32_t       viol = 0;
                // int16 -> fp16 (I / 512) in place, thread per row
#pragma unroll
                for(int ch = 0; ch < 8; ch++) {
                    uint4    *cp = reinterpret_cast<uint4 *>(my_rowp + ((ch ^ rsw) << 4));
                    uint4     w = *cp;
                    uint32_t *pw = &w.x;
#pragma unroll
                    for(int i = 0; i < 4; i++) {
                        if(kCheck) viol |= pw[i] ^ (pw[i] << 1); // bits 15..10 of each half all equal <=> in [-1024, 1023]
                        uint32_t x;
                        asm("lop3.b32 %0, %1, 0x07FF07FF, %2, 0x6A;" : "=r"(x) : "r"(pw[i]), "r"(kx)); // (w & mask) ^ kx
                        asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(pw[i]) : "r"(x), "r"(0x78007800u), "r"(0xC000C000u));
                    }
                    *cp = w;
                }
                unsigned long long *addr = waddr + st * 32;
                const bool          mine = addr[lane] != 0ull;
                const bool          bad = kCheck && mine && (viol & 0xF800F800u) != 0;
                if(bad) { // leave this block to the fp32 kernel: one bit per (list slot, image)
                    const int slot = blockIdx.x + j * gridDim.x;
                    atomicOr(p.redo_mask + (size_t)(slot >> 5) * p.n + cur.img, 1u << (slot & 31));
                    atomicAdd(p.redo_count, 1u);
                    addr[lane] = 0ull; // ... and do not store it
                }
                fence_async_smem(); // the rows, as the tensor core will read them
                tc_fence_before();  // ... and my tcgen05.ld of the previous unit
                asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");
                if(wq == 0 && lane == 0) {
                    tc_fence_after();
                    const uint32_t bbuf = base32 + L::kB + (j & 1) * L::kBBuf;
                    const uint64_t ad = umma_desc_sw128(base32 + L::kStage + (grp * kOpStages + st) * kOpStageBytes);
                    const uint64_t bh = umma_desc_sw128(bbuf);
#pragma unroll
                    for(int ks = 0; ks < 4; ks++) umma_f16(tmem_g, ad + 2 * ks, bh + 2 * ks, umma_idesc_f16(128, 128), ks > 0);
#pragma unroll
                    for(int pc = 1; pc < NP; pc++) {
                        const uint64_t bl = umma_desc_sw128(bbuf + (pc + 1) * kOpPieceBytes);
#pragma unroll
                        for(int ks = 0; ks < 4; ks++) umma_f16(tmem_g, ad + 2 * ks, bl + 2 * ks, umma_idesc_f16(128, 64), 1);
                    }
                    umma_commit(bar_mma);
                }
                __syncwarp();
                mbar_wait(bar_mma, mma_phase);
                mma_phase ^= 1;
                tc_fence_after();
                if(k == bpg - 1) handoff(j); // every UMMA of this group on item j has completed

                const float  ascale = (float)(1 << (9 - (int)sMisc[4 + c]));
                const float *sK = reinterpret_cast<const float *>(base + L::kK + (j & (kOpKRing - 1)) * 256);
                const float *rqc = sRq + c * 64;
#pragma unroll
                for(int h = 0; h < 4; h++) {
                    float y[16], iq[16];
                    tmem_ld16(taddr + 16 * h, y);
                    tmem_ld16(taddr + 64 + 16 * h, iq);
                    tmem_wait_ld();
#pragma unroll
                    for(int r2 = 0; r2 < 2; r2++) {
                        const int r = 2 * h + r2;
                        uint32_t  o[4];
#pragma unroll
                        for(int q4 = 0; q4 < 2; q4++) {
                            const float4 kk = *reinterpret_cast<const float4 *>(sK + r * 8 + 4 * q4);
                            const float4 rr = *reinterpret_cast<const float4 *>(rqc + r * 8 + 4 * q4);
                            const float *yy = y + 8 * r2 + 4 * q4, *ii = iq + 8 * r2 + 4 * q4;
                            o[2 * q4] = requant_pair_acc(f2(yy[0], yy[1]), ascale, f2(kk.x, kk.y), f2(ii[0], ii[1]), f2(rr.x, rr.y));
                            o[2 * q4 + 1] = requant_pair_acc(f2(yy[2], yy[3]), ascale, f2(kk.z, kk.w), f2(ii[2], ii[3]), f2(rr.z, rr.w));
                        }
                        if(mine && !bad) *reinterpret_cast<uint4 *>(my_rowp + ((r ^ rsw) << 4)) = make_uint4(o[0], o[1], o[2], o[3]);
                    }
                }
                __syncwarp();
                // coalesced write-back: 8 lanes per row, lane -> chunk (lane & 7) of rows (lane >> 3) + 4j
                {
                    const unsigned char      *src = wstage0 + st * kOpStageBytes;
                    const unsigned long long *ap = addr + (lane >> 3);
                    const unsigned            coff = (lane & 7) * 16;
                    unsigned long long        b[8];
                    uint4                     v[8];
#pragma unroll
                    for(int jj = 0; jj < 8; jj++) {
                        b[jj] = ap[4 * jj];
                        v[jj] = *reinterpret_cast<const uint4 *>(src + jj * 512 + ((jj & 1) ? cp_off1 : cp_off0));
                    }
#pragma unroll
                    for(int jj = 0; jj < 8; jj++)
                        if(b[jj]) __stcs(reinterpret_cast<uint4 *>(b[jj] + coff), v[jj]);
                }
            }
            else {
                // skipped item (padding slot, component not served): the group still passes it together, so that no warp
                // is ever two phases behind the buffer barrier
                asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");
                if(k == bpg - 1) handoff(j);
            }
            __syncwarp();
            cur = nxt;
        }
        cp_async_wait<0>();
    }
    // every warp is done with its tensor-memory lanes before warp 0 returns the allocation
    tc_fence_before();
    __syncthreads();
    if(widx == 0) tmem_dealloc<512>(sMisc[0]);
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
    const size_t oD = take(MJX_MAX_COMPONENTS * kOpPieceBytes);
    const size_t oR = take(MJX_MAX_COMPONENTS * 256);
    const size_t oKey = take(MJX_MAX_COMPONENTS * 128);
    const size_t oI = take(MJX_MAX_COMPONENTS * 4);
    const size_t oRb = take(MJX_MAX_COMPONENTS * 4);
    if(v) {
        unsigned char *b = v->B; // the caller put the slab's base here
        v->B = b + oB;
        v->K = reinterpret_cast<float *>(b + oK);
        v->diag = b + oD;
        v->rq = reinterpret_cast<float *>(b + oR);
        v->key = reinterpret_cast<uint16_t *>(b + oKey);
        v->info = reinterpret_cast<int *>(b + oI);
        v->rebuild = reinterpret_cast<int *>(b + oRb);
        v->np = np;
    }
    return off;
}

template <int NP, int G>
static cudaError_t launch_op_kernel(cudaStream_t s, const OpParams &p, int sms, bool check, bool *attr_set) {
    using L = OpSmem<NP, G>;
    cudaError_t e;
    if(!*attr_set) {
        if((e = cudaFuncSetAttribute(k2_generic_op_kernel<NP, G, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kBytes)) != cudaSuccess) return e;
        if((e = cudaFuncSetAttribute(k2_generic_op_kernel<NP, G, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kBytes)) != cudaSuccess) return e;
        *attr_set = true;
    }
    const int ctas = p.drop.n_generic < sms ? p.drop.n_generic : sms;
    if(check) k2_generic_op_kernel<NP, G, true><<<ctas, G * 128, L::kBytes, s>>>(p);
    else k2_generic_op_kernel<NP, G, false><<<ctas, G * 128, L::kBytes, s>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_k2_generic_op(cudaStream_t s, const OpParams &p, int sm_count, bool check, bool *attr_set, int *launches) {
    cudaError_t e;
    const int   sms = sm_count > 0 ? sm_count : 148;
    k2_op_prepare_kernel<<<(p.n + 127) / 128, 128, 0, s>>>(p);
    if((e = cudaGetLastError()) != cudaSuccess) return e;
    if(p.op.np == 3) k2_op_build_kernel<3><<<p.drop.n_generic, 256, 0, s>>>(p);
    else k2_op_build_kernel<2><<<p.drop.n_generic, 256, 0, s>>>(p);
    if((e = cudaGetLastError()) != cudaSuccess) return e;
    if(p.op.np == 3) e = launch_op_kernel<3, kOpGroups>(s, p, sms, check, attr_set);
    else e = launch_op_kernel<2, kOpGroups>(s, p, sms, check, attr_set + 1);
    if(e != cudaSuccess) return e;
    if(launches) *launches += 3;
    return cudaSuccess;
}

} // namespace mjx
