// k2_generic_tc.cu -- K2 for class G blocks with the INVERSE half on the tensor cores (tcgen05 + TMEM).
// Replaces, like k2_generic_kernel, mj_compose_with_mask + mj_convolve for blocks with a non-uniform mask
// (reference: src/compose.c:237-342, src/convolve.c:29-1099).
//
// Why: the fp32 G kernel is bound by the fp32 pipe and by int16 -> fp32 conversions, not by HBM.  Of its
// two 2-D transforms, the inverse one acts on an operand that is an EXACT small integer: the quantised coefficients I.
//     i = IDCT2(I o q)  =  I[1 x 64] . M_q[64 x 64],     M_q = diag(q) (C (x) C)         (per quantisation table)
// so a batch of 128 blocks is one [128 x 64] x [64 x 128] UMMA (kind::f16, fp32 accumulation in tensor memory):
//   * A operand = the raw int16 rows, converted IN PLACE to fp16 with one LOP3 + one HFMA2 per coefficient PAIR:
//     for I in [-1024, 1023] (every coefficient a baseline JPEG can carry: DC 11 bits, AC 10 bits) the 11-bit field
//     (I + 1024) is the fp16 bit pattern of (I + 1024) * 2^-24, and fma(that, 2^15, -2) = I / 512 exactly.  The rows are
//     staged by cp.async straight into the canonical K-major SWIZZLE_128B layout, so the conversion is 8 LDS.128 +
//     8 STS.128 per block and the tensor core reads what the copy engine wrote.
//   * B operand, built in shared memory per quantisation table: columns 0..63 = S * M_q split into two fp16 pieces
//     (hi + lo, 22 significant bits; S = 512 for 8-bit tables so the products come out as pixels), columns 64..127 =
//     diag(q): the accumulator then also holds I*q / 512 EXACTLY, which is what the requantisation needs -- the second
//     int16 -> fp32 conversion pass of the fp32 kernel disappears as well.
//   * each thread reads the 64 pixels and the 64 products of ITS block from its own TMEM lane (tcgen05.ld.32x32b) and
//     carries on as before in packed fp32: y = A o (d - i) as one FFMA2 against the per-dropon constants A and A o d,
//     forward AAN transform, requantisation (mjx_math.cuh), results back over the staged rows, coalesced stores.
// Measured against double precision (profiles/microbench/tc_idct_ubench.cu): pixel error rms 8.7e-6 / max 5.8e-5 for
// JPEG-like blocks, the fp32 AAN path has 5.1e-6 / 3.1e-5; I*q exact.
//
// One CTA of 12 warps per SM, three groups of four warps; a group = one UMMA per iteration (M = 128 rows = 4 images x 32
// list entries of the tile).  The three groups share the tile's constants and B and run out of phase, which is what hides
// the issue -> commit -> mbarrier latency of the tensor pipe (the same job three warps per scheduler do for each other
// in the fp32 kernel).  Work item = (tile of 32 list entries, chunk of <= 96 images); inside an item the images are
// grouped by quantisation table (any order, any mix -- one B build per distinct table and item).
//
// Contract: coefficients in [-1024, 1023].  With kCheck the kernel verifies that per block (2 instructions per pair) and
// leaves the (tile, image) pairs that fail to the fp32 kernel (redo bitmap); without it the caller vouches for the range.
#include <cuda_fp16.h>

#include "k2_common.cuh"

namespace mjx {

static constexpr int kTcGroups = 3;
static constexpr int kTcWarps = 4 * kTcGroups;
static constexpr int kTcThreads = 32 * kTcWarps;
static constexpr int kTcStages = 2;
static constexpr int kTcStageBytes = 128 * 128; // one group's 128 rows x 64 fp16 / int16
static constexpr int kTcChunk = 96;             // images per work item (three 32-bit class masks)
static constexpr int kTcCols = 128;             // TMEM columns per group: 64 pixels + 64 products

// shared memory map, offsets from a 1024-byte aligned base (SWIZZLE_128B atoms are 1024 bytes)
static constexpr int kOffStage = 0;                                                  // [group][stage] 16 KB
static constexpr int kOffBh = kOffStage + kTcGroups * kTcStages * kTcStageBytes;     // B hi piece, 128 rows
static constexpr int kOffBl = kOffBh + 128 * 128;                                    // B lo piece, 128 rows
static constexpr int kOffTile = kOffBl + 128 * 128;                                  // A, then A o d: 32 x 272 B each
static constexpr int kOffAddr = kOffTile + kTileBytes;                               // [warp][stage][32] global addresses
static constexpr int kOffRq = kOffAddr + kTcWarps * kTcStages * 256;                 // biased reciprocals, 64 floats
static constexpr int kOffQcur = kOffRq + 256;                                        // table B was built from (128 B)
static constexpr int kOffQrep = kOffQcur + 128;                                      // table of the class being formed
static constexpr int kOffDct = kOffQrep + 128;                                       // C[k][n], 64 doubles
static constexpr int kOffBar = kOffDct + 512;                                        // one mbarrier per group
static constexpr int kOffMisc = kOffBar + 32;                                        // tmem base, item, class masks, scale
static constexpr int kTcSmemBytes = kOffMisc + 64 + 1024;                            // + alignment slack

static __constant__ double c_dct8[64] = {
    0.35355339059327379, 0.35355339059327379, 0.35355339059327379, 0.35355339059327379, 0.35355339059327379, 0.35355339059327379, 0.35355339059327379, 0.35355339059327379,
    0.49039264020161522, 0.41573480615127262, 0.27778511650980114, 0.097545161008064166, -0.097545161008064166, -0.27778511650980114, -0.41573480615127262, -0.49039264020161522,
    0.46193976625564337, 0.19134171618254492, -0.19134171618254492, -0.46193976625564337, -0.46193976625564337, -0.19134171618254492, 0.19134171618254492, 0.46193976625564337,
    0.41573480615127262, -0.097545161008064166, -0.49039264020161522, -0.27778511650980114, 0.27778511650980114, 0.49039264020161522, 0.097545161008064166, -0.41573480615127262,
    0.35355339059327379, -0.35355339059327379, -0.35355339059327379, 0.35355339059327379, 0.35355339059327379, -0.35355339059327379, -0.35355339059327379, 0.35355339059327379,
    0.27778511650980114, -0.49039264020161522, 0.097545161008064166, 0.41573480615127262, -0.41573480615127262, -0.097545161008064166, 0.49039264020161522, -0.27778511650980114,
    0.19134171618254492, -0.46193976625564337, 0.46193976625564337, -0.19134171618254492, -0.19134171618254492, 0.46193976625564337, -0.46193976625564337, 0.19134171618254492,
    0.097545161008064166, -0.27778511650980114, 0.41573480615127262, -0.49039264020161522, 0.49039264020161522, -0.41573480615127262, 0.27778511650980114, -0.097545161008064166};

// ---- tcgen05 / mbarrier plumbing ---------------------------------------------------------------------------------------
// instruction descriptor, kind::f16: D fp32 (bits 4-5 = 1), A and B fp16 (0) and K-major (bits 15, 16 = 0), N >> 3 at bit 17,
// M >> 4 at bit 24
static constexpr uint32_t kIdesc = (1u << 4) | ((128u >> 3) << 17) | ((128u >> 4) << 24);

// shared-memory matrix descriptor: K-major, SWIZZLE_128B (one 128-byte row per matrix row, 8-row atoms 1024 bytes apart)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)(1024u >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d), "l"(adesc),
                 "l"(bdesc), "r"(kIdesc), "r"(accumulate)
                 : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// try_wait suspends the thread in hardware for a bounded time per attempt; a tensor-pipe batch takes ~1 us.  The attempt
// count is bounded so that a descriptor bug traps (the launch fails with an error) instead of hanging the device.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    for(int spin = 0; spin < (1 << 24); spin++) {
        uint32_t ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if(ok) return;
    }
    __trap();
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float *v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, "
        "%24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]), "=f"(v[8]), "=f"(v[9]), "=f"(v[10]), "=f"(v[11]),
          "=f"(v[12]), "=f"(v[13]), "=f"(v[14]), "=f"(v[15]), "=f"(v[16]), "=f"(v[17]), "=f"(v[18]), "=f"(v[19]), "=f"(v[20]), "=f"(v[21]), "=f"(v[22]),
          "=f"(v[23]), "=f"(v[24]), "=f"(v[25]), "=f"(v[26]), "=f"(v[27]), "=f"(v[28]), "=f"(v[29]), "=f"(v[30]), "=f"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float *v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// position of the o-th (0-based) set bit of a 96-bit mask (the caller guarantees it exists)
__device__ __forceinline__ int nth_set96(uint32_t m0, uint32_t m1, uint32_t m2, int o) {
    const int c0 = __popc(m0), c1 = __popc(m1);
    if(o < c0) return (int)__fns(m0, 0, o + 1);
    o -= c0;
    if(o < c1) return 32 + (int)__fns(m1, 0, o + 1);
    return 64 + (int)__fns(m2, 0, o - c1 + 1);
}

// requantisation of one pair with the exact product I*q / 512 from the accumulator (mjx_math.cuh: requant_pair)
__device__ __forceinline__ uint32_t requant_pair_iq(F2 y, F2 f, F2 iqs, F2 rq) {
    const F2 sm = signed_magic2(y);
    const F2 t = sub2(fma2_rz(y, f, sm), sm);
    const F2 a = fma2(iqs, bc2(512.0f), t);
    const F2 sa = signed_magic2(a);
    const F2 m = fma2_rz(a, rq, sa);
    const F2 o = add2(m, sub2(bc2(12582912.0f), sa));
    return __byte_perm(__float_as_uint(o.x), __float_as_uint(o.y), 0x5410);
}

__device__ __forceinline__ uint32_t pack_half2(float lo, float hi) {
    const __half2 h = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<const uint32_t *>(&h);
}

template <bool kCheck>
__global__ void __launch_bounds__(kTcThreads, 1) k2_generic_tc_kernel(const FastParams p) {
    extern __shared__ unsigned char smem_dyn[];
    unsigned char *base = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~static_cast<uintptr_t>(1023));
    const int      lane = threadIdx.x & 31, widx = threadIdx.x >> 5, grp = widx >> 2, wq = widx & 3;

    unsigned char      *sBh = base + kOffBh, *sBl = base + kOffBl;
    const unsigned char *myA = base + kOffTile + lane * kF32Stride, *myAd = myA + kTileHalf;
    float              *sRq = reinterpret_cast<float *>(base + kOffRq);
    uint32_t           *sQcur = reinterpret_cast<uint32_t *>(base + kOffQcur), *sQrep = reinterpret_cast<uint32_t *>(base + kOffQrep);
    double             *sDct = reinterpret_cast<double *>(base + kOffDct);
    uint32_t           *sMisc = reinterpret_cast<uint32_t *>(base + kOffMisc);
    const uint32_t      base32 = smem_u32(base);
    const uint32_t      bar = base32 + kOffBar + 8 * grp;

    // ---- one-time set-up: tensor memory (all 512 columns: one CTA per SM), mbarriers, the DCT matrix ----
    if(widx == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sMisc[0])), "n"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if(threadIdx.x == 32) {
#pragma unroll
        for(int g = 0; g < kTcGroups; g++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(base32 + kOffBar + 8 * g) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if(threadIdx.x >= 64 && threadIdx.x < 128) sDct[threadIdx.x - 64] = c_dct8[threadIdx.x - 64];
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_g = sMisc[0] + grp * kTcCols;                  // this group's accumulator
    const uint32_t taddr = tmem_g + ((uint32_t)(wq * 32) << 16);        // this warp's lanes of it
    uint32_t       phase = 0;                                           // parity of the group's mbarrier
    bool           have_b = false;
    float          ascale = 1.0f;                                       // pixels = accumulator * ascale (1 for 8-bit tables)

    const int ntiles = p.drop.n_generic >> 5; // the list is padded to whole single-component tiles
    const int nchunks = (p.n + kTcChunk - 1) / kTcChunk;
    const int nitems = ntiles * nchunks;

    // per-lane constants of the staging layout.  Row r of the warp's 32 lies at r * 128; chunk c of row r at (c ^ (r & 7)) * 16.
    unsigned char *wstage0 = base + kOffStage + (grp * kTcStages) * kTcStageBytes + wq * 4096; // stage 0 of this warp
    const uint32_t wstage0_32 = smem_u32(wstage0);
    //   thread-per-row view (conversion, results): my row = lane
    const int      rsw = lane & 7;
    //   8-lanes-per-row view (cp.async in, coalesced stores out): lane -> chunk (lane & 7) of rows (lane >> 3) + 4 j
    const uint32_t cp_off0 = (uint32_t)(lane >> 3) * 128 + (uint32_t)(((lane & 7) ^ (lane >> 3)) << 4);       // j even
    const uint32_t cp_off1 = (uint32_t)(lane >> 3) * 128 + (uint32_t)(((lane & 7) ^ ((lane >> 3) + 4)) << 4); // j odd
    unsigned long long *waddr = reinterpret_cast<unsigned long long *>(base + kOffAddr) + widx * kTcStages * 32;

    uint32_t kx; // (w & 0x07FF07FF) ^ kx as ONE LOP3: the constant must live in a register
    asm volatile("mov.u32 %0, 0x04000400;" : "=r"(kx));

    int  cur_tile = -1, tile_c = 0;
    int  my_row = 0, my_col = 0;
    bool my_valid = false;

    for(;;) {
        __syncthreads(); // every warp is done with the previous item (tile constants, B, sMisc)
        if(threadIdx.x == 0) sMisc[1] = atomicAdd(p.counter, 1u);
        __syncthreads();
        const int item = (int)sMisc[1];
        if(item >= nitems) break;
        const int tile = item / nchunks, chunk = item - tile * nchunks;
        if(tile != cur_tile) {
            cur_tile = tile;
            const uint32_t e = __ldg(p.drop.list_generic + tile * 32 + lane);
            my_valid = e != 0xffffffffu;
            tile_c = 0; // tiles hold one component only (the list is padded per component)
#pragma unroll
            for(int c = 1; c < MJX_MAX_COMPONENTS; c++)
                if(c < p.drop.ncomp && tile >= p.drop.gtile_start[c]) tile_c = c;
            if(my_valid) {
                const DropComp &dc = p.drop.comp[tile_c];
                my_row = p.block_y * dc.vs + entry_row(e);
                my_col = p.block_x * dc.hs + entry_col(e);
            }
            // A and A o d of the tile: 2 x 512 chunks of 16 B; chunk i -> block i >> 4, chunk i & 15 of the padded row
            const float *ga = p.drop.gA + (size_t)tile * 32 * 64, *gd = p.drop.gAd + (size_t)tile * 32 * 64;
            for(int i = threadIdx.x; i < 512; i += kTcThreads) {
                const uint32_t dst = base32 + kOffTile + (i >> 4) * kF32Stride + (i & 15) * 16;
                cp_async16(dst, ga + i * 4);
                cp_async16(dst + kTileHalf, gd + i * 4);
            }
            cp_async_commit();
            cp_async_wait<0>();
        }
        const int i0 = chunk * kTcChunk, cnt = min(p.n - i0, kTcChunk);
        uint32_t  rem0 = cnt >= 32 ? 0xffffffffu : (1u << cnt) - 1u;
        uint32_t  rem1 = cnt >= 64 ? 0xffffffffu : (cnt > 32 ? (1u << (cnt - 32)) - 1u : 0u);
        uint32_t  rem2 = cnt >= 96 ? 0xffffffffu : (cnt > 64 ? (1u << (cnt - 64)) - 1u : 0u);

        // ---- the chunk's images, one class (= one quantisation table of this component) at a time ----
        while(rem0 | rem1 | rem2) {
            const int rep = rem0 ? __ffs(rem0) - 1 : (rem1 ? 31 + __ffs(rem1) : 63 + __ffs(rem2));
            if(threadIdx.x < 8)
                reinterpret_cast<uint4 *>(sQrep)[threadIdx.x] = __ldg(reinterpret_cast<const uint4 *>(&p.items[i0 + rep].q[tile_c][0]) + threadIdx.x);
            if(threadIdx.x >= 32 && threadIdx.x < 35) sMisc[2 + threadIdx.x - 32] = 0;
            __syncthreads();
            // members: images still to do whose table equals the representative's (8 threads per image, 48 images per pass)
            for(int ps = 0; ps * 48 < cnt; ps++) {
                const int  j = ps * 48 + (threadIdx.x >> 3);
                const bool cand = j < cnt && (((j < 32 ? rem0 : (j < 64 ? rem1 : rem2)) >> (j & 31)) & 1u);
                bool       eq = false;
                if(cand) {
                    const uint4 a = __ldg(reinterpret_cast<const uint4 *>(&p.items[i0 + j].q[tile_c][0]) + (threadIdx.x & 7));
                    const uint4 b = reinterpret_cast<const uint4 *>(sQrep)[threadIdx.x & 7];
                    eq = a.x == b.x && a.y == b.y && a.z == b.z && a.w == b.w;
                }
                const unsigned bal = __ballot_sync(0xffffffffu, eq);
                if(lane == 0) {
#pragma unroll
                    for(int g4 = 0; g4 < 4; g4++)
                        if(((bal >> (8 * g4)) & 0xffu) == 0xffu) {
                            const int jj = ps * 48 + widx * 4 + g4;
                            atomicOr(&sMisc[2 + (jj >> 5)], 1u << (jj & 31));
                        }
                }
            }
            bool differs = !have_b;
            if(threadIdx.x < 32) differs = differs || sQrep[threadIdx.x] != sQcur[threadIdx.x];
            const int rebuild = __syncthreads_or(differs ? 1 : 0); // also publishes the class masks
            const uint32_t cls0 = sMisc[2], cls1 = sMisc[3], cls2 = sMisc[4];
            rem0 &= ~cls0, rem1 &= ~cls1, rem2 &= ~cls2;

            if(rebuild) {
                // scale S: the largest power of two <= 512 with 0.25 * qmax * S <= 32768 (fp16 range); 512 for every 8-bit table
                if(widx == 0) {
                    const uint32_t w = sQrep[lane];
                    unsigned       qm = max(max(w & 0xffffu, w >> 16), 1u);
                    qm = __reduce_max_sync(0xffffffffu, qm);
                    const int lg = qm > 1 ? 32 - __clz((int)qm - 1) : 0; // ceil(log2(qmax))
                    const int sh = min(9, 17 - lg);
                    if(lane == 0) sMisc[5] = (uint32_t)sh;
                    sQcur[lane] = w;
                }
                __syncthreads();
                const int    sh = (int)sMisc[5];
                const double S = (double)(1 << sh);
                ascale = (float)(1 << (9 - sh));
                const uint16_t *q16 = reinterpret_cast<const uint16_t *>(sQrep);
                // pixel columns: row j of B <-> pixel (y = 2i + h, x = k) with j = 2 (8 i + k) + h (the Q pairing); 8 coefficients per chunk
                for(int ci = threadIdx.x; ci < 512; ci += kTcThreads) {
                    const int    j = ci >> 3, v = ci & 7; // chunk v holds coefficients (v, u = 0..7)
                    const int    y = 2 * (j >> 4) + (j & 1), x = (j >> 1) & 7;
                    const double cy = S * sDct[v * 8 + y];
                    uint32_t     hi[4], lo[4];
#pragma unroll
                    for(int u2 = 0; u2 < 4; u2++) {
                        float fh[2], fl[2];
#pragma unroll
                        for(int e2 = 0; e2 < 2; e2++) {
                            const int    u = 2 * u2 + e2;
                            const double m = cy * sDct[u * 8 + x] * (double)max((int)q16[v * 8 + u], 1);
                            const __half h = __float2half_rn((float)m);
                            fh[e2] = __half2float(h);
                            fl[e2] = (float)(m - (double)fh[e2]);
                        }
                        hi[u2] = pack_half2(fh[0], fh[1]);
                        lo[u2] = pack_half2(fl[0], fl[1]);
                    }
                    const int off = (j >> 3) * 1024 + (j & 7) * 128 + ((v ^ (j & 7)) << 4);
                    *reinterpret_cast<uint4 *>(sBh + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                    *reinterpret_cast<uint4 *>(sBl + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                }
                // product columns: row 64 + c of B = q[c] at k = c (hi + lo pieces: exact for every 16-bit q)
                for(int ci = threadIdx.x; ci < 512; ci += kTcThreads) {
                    const int c = ci >> 3, v = ci & 7, row = 64 + c;
                    uint32_t  hi[4] = {0u, 0u, 0u, 0u}, lo[4] = {0u, 0u, 0u, 0u};
                    if(v == (c >> 3)) {
                        const float  qf = (float)max((int)q16[c], 1);
                        const __half h = __float2half_rn(qf);
                        const float  r = qf - __half2float(h);
                        const int    pos = c & 7;
                        hi[pos >> 1] = (pos & 1) ? pack_half2(0.f, __half2float(h)) : pack_half2(__half2float(h), 0.f);
                        lo[pos >> 1] = (pos & 1) ? pack_half2(0.f, r) : pack_half2(r, 0.f);
                    }
                    const int off = (row >> 3) * 1024 + (row & 7) * 128 + ((v ^ (row & 7)) << 4);
                    *reinterpret_cast<uint4 *>(sBh + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                    *reinterpret_cast<uint4 *>(sBl + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                }
                if(threadIdx.x < 64) sRq[threadIdx.x] = quant_rcp_fast((float)max((int)q16[threadIdx.x], 1));
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncthreads();
                have_b = true;
            }

            // ---- the class: `total` images, 12 per iteration (one per warp) ----
            const int total = __popc(cls0) + __popc(cls1) + __popc(cls2);
            const int iters = (total + kTcWarps - 1) / kTcWarps;
            // lane l holds plane pointer / stride / rows of this warp's image of iteration l
            unsigned long long d_plane = 0;
            int                d_stride = 0, d_rows = 0;
            {
                const int o = lane * kTcWarps + widx;
                if(o < total) {
                    const mjx_image_desc_t &im = p.items[i0 + nth_set96(cls0, cls1, cls2, o)];
                    d_plane = im.plane[tile_c];
                    d_stride = im.stride_blocks[tile_c];
                    d_rows = im.rows[tile_c];
                }
            }
            // issue the loads of this warp's image of iteration `it` into stage `st`
            auto prefetch = [&](int it, int st) {
                const unsigned long long plane = __shfl_sync(0xffffffffu, d_plane, it);
                const int                stride = __shfl_sync(0xffffffffu, d_stride, it), rows = __shfl_sync(0xffffffffu, d_rows, it);
                unsigned long long      *addr = waddr + st * 32;
                unsigned long long       a = 0;
                if(my_valid && plane != 0 && my_row < rows && my_col < stride) a = plane + ((unsigned long long)my_row * stride + my_col) * 128ull;
                addr[lane] = a;
                const bool all_there = __all_sync(0xffffffffu, a != 0);
                __syncwarp();
                if(plane == 0) return; // no image for this warp in this iteration
                const uint32_t            dst = wstage0_32 + st * kTcStageBytes;
                const unsigned long long *ap = addr + (lane >> 3);
                const unsigned            coff = (lane & 7) * 16;
                unsigned long long        b[8];
#pragma unroll
                for(int j = 0; j < 8; j++) b[j] = ap[4 * j];
                if(all_there) {
#pragma unroll
                    for(int j = 0; j < 8; j++) cp_async16(dst + j * 512 + ((j & 1) ? cp_off1 : cp_off0), reinterpret_cast<const void *>(b[j] + coff));
                }
                else {
#pragma unroll
                    for(int j = 0; j < 8; j++)
                        cp_async16(dst + j * 512 + ((j & 1) ? cp_off1 : cp_off0),
                                   reinterpret_cast<const void *>((b[j] ? b[j] : (unsigned long long)(uintptr_t)p.items) + coff), b[j] ? 16u : 0u);
                }
            };

            prefetch(0, 0);
            cp_async_commit();
            for(int it = 0; it < iters; it++) {
                const int  st = it & 1;
                const bool have = it * kTcWarps + widx < total; // warp-uniform
                if(it + 1 < iters) prefetch(it + 1, st ^ 1);
                cp_async_commit();
                cp_async_wait<1>(); // everything but the newest group: this iteration's rows have landed
                __syncwarp();

                unsigned char *my_rowp = wstage0 + st * kTcStageBytes + lane * 128;
                uint32_t       viol = 0;
                if(have) {
                    // int16 -> fp16 (I / 512) in place, thread per row
#pragma unroll
                    for(int c = 0; c < 8; c++) {
                        uint4    *cp = reinterpret_cast<uint4 *>(my_rowp + ((c ^ rsw) << 4));
                        uint4     w = *cp;
                        uint32_t *pw = &w.x;
#pragma unroll
                        for(int k = 0; k < 4; k++) {
                            if(kCheck) viol |= pw[k] ^ (pw[k] << 1); // bits 15..10 of each half all equal <=> in [-1024, 1023]
                            uint32_t u;
                            asm("lop3.b32 %0, %1, 0x07FF07FF, %2, 0x6A;" : "=r"(u) : "r"(pw[k]), "r"(kx)); // (w & mask) ^ kx
                            asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(pw[k]) : "r"(u), "r"(0x78007800u), "r"(0xC000C000u));
                        }
                        *cp = w;
                    }
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); // the rows, as the tensor core will read them
                tc_fence_before();                                            // ... and my tcgen05.ld of the previous iteration
                asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");
                if(wq == 0 && lane == 0) {
                    tc_fence_after();
                    const uint64_t ad = umma_desc(base32 + kOffStage + (grp * kTcStages + st) * kTcStageBytes);
                    const uint64_t bh = umma_desc(base32 + kOffBh), bl = umma_desc(base32 + kOffBl);
#pragma unroll
                    for(int k = 0; k < 4; k++) umma_f16(tmem_g, ad + 2 * k, bh + 2 * k, k > 0);
#pragma unroll
                    for(int k = 0; k < 4; k++) umma_f16(tmem_g, ad + 2 * k, bl + 2 * k, 1);
                    umma_commit(bar);
                }
                __syncwarp();
                mbar_wait(bar, phase);
                phase ^= 1;
                tc_fence_after();

                bool redo = false;
                if(kCheck) redo = __any_sync(0xffffffffu, (viol & 0xF800F800u) != 0);
                if(have && !redo) {
                    F2 x[32], y[32];
                    {
                        float v[64];
                        tmem_ld32(taddr, v);
                        tmem_ld32(taddr + 32, v + 32);
                        tmem_wait_ld();
                        if(ascale != 1.0f) {
#pragma unroll
                            for(int i = 0; i < 64; i++) v[i] *= ascale;
                        }
                        // y = A o (d - i) = (A o d) - A * i, Q-paired like A
#pragma unroll
                        for(int c = 0; c < 16; c++) {
                            const float4 a = *reinterpret_cast<const float4 *>(myA + c * 16);
                            const float4 d = *reinterpret_cast<const float4 *>(myAd + c * 16);
                            y[2 * c] = fma2(f2(-v[4 * c], -v[4 * c + 1]), f2(a.x, a.y), f2(d.x, d.y));
                            y[2 * c + 1] = fma2(f2(-v[4 * c + 2], -v[4 * c + 3]), f2(a.z, a.w), f2(d.z, d.w));
                        }
                    }
#pragma unroll
                    for(int i = 0; i < 4; i++) fdct8p_rowpairs_to_cols(y, x, i);
#pragma unroll
                    for(int j = 0; j < 4; j++) fdct8p<4>(x + j);
#pragma unroll
                    for(int r = 0; r < 8; r++) {
                        float iq[8];
                        tmem_ld8(taddr + 64 + 8 * r, iq);
                        const float4 r0 = *reinterpret_cast<const float4 *>(sRq + r * 8);
                        const float4 r1 = *reinterpret_cast<const float4 *>(sRq + r * 8 + 4);
                        tmem_wait_ld();
                        uint4 o;
                        o.x = requant_pair_iq(x[4 * r + 0], c_fwd2.v[4 * r + 0], f2(iq[0], iq[1]), f2(r0.x, r0.y));
                        o.y = requant_pair_iq(x[4 * r + 1], c_fwd2.v[4 * r + 1], f2(iq[2], iq[3]), f2(r0.z, r0.w));
                        o.z = requant_pair_iq(x[4 * r + 2], c_fwd2.v[4 * r + 2], f2(iq[4], iq[5]), f2(r1.x, r1.y));
                        o.w = requant_pair_iq(x[4 * r + 3], c_fwd2.v[4 * r + 3], f2(iq[6], iq[7]), f2(r1.z, r1.w));
                        *reinterpret_cast<uint4 *>(my_rowp + ((r ^ rsw) << 4)) = o;
                    }
                    __syncwarp();
                    // coalesced write-back: 8 lanes per row, lane -> chunk (lane & 7) of rows (lane >> 3) + 4j
                    {
                        const unsigned char      *src = wstage0 + st * kTcStageBytes;
                        const unsigned long long *ap = waddr + st * 32 + (lane >> 3);
                        const unsigned            coff = (lane & 7) * 16;
                        unsigned long long        b[8];
                        uint4                     v[8];
#pragma unroll
                        for(int j = 0; j < 8; j++) {
                            b[j] = ap[4 * j];
                            v[j] = *reinterpret_cast<const uint4 *>(src + j * 512 + ((j & 1) ? cp_off1 : cp_off0));
                        }
#pragma unroll
                        for(int j = 0; j < 8; j++)
                            if(b[j]) __stcs(reinterpret_cast<uint4 *>(b[j] + coff), v[j]);
                    }
                }
                else if(kCheck && have && redo) {
                    // leave this (tile, image) to the fp32 kernel: one bit per pair
                    if(lane == 0) {
                        const int       img = i0 + nth_set96(cls0, cls1, cls2, it * kTcWarps + widx);
                        const long long bit = (long long)tile * p.n + img;
                        atomicOr(p.redo_bits + (bit >> 5), 1u << (bit & 31));
                        atomicAdd(p.redo_count, 1u);
                    }
                }
                __syncwarp();
            }
            cp_async_wait<0>();
            __syncthreads(); // the class is done: sQrep / sMisc / B may change
        }
    }
    // every warp is done with its tensor-memory lanes before warp 0 returns the allocation
    tc_fence_before();
    __syncthreads();
    if(widx == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(sMisc[0]), "n"(512) : "memory");
}

cudaError_t launch_k2_generic_tc(cudaStream_t s, const FastParams &p, int sm_count, bool check, bool *attr_set) {
    cudaError_t e;
    if(!*attr_set) {
        if((e = cudaFuncSetAttribute(k2_generic_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmemBytes)) != cudaSuccess) return e;
        if((e = cudaFuncSetAttribute(k2_generic_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmemBytes)) != cudaSuccess) return e;
        *attr_set = true;
    }
    const long long nitems = (long long)(p.drop.n_generic / 32) * ((p.n + kTcChunk - 1) / kTcChunk);
    if(nitems > 0x7fffffffLL) return cudaErrorInvalidValue;
    const int sms = sm_count > 0 ? sm_count : 148;
    const int ctas = nitems < sms ? (int)nitems : sms;
    if(check) k2_generic_tc_kernel<true><<<ctas, kTcThreads, kTcSmemBytes, s>>>(p);
    else k2_generic_tc_kernel<false><<<ctas, kTcThreads, kTcSmemBytes, s>>>(p);
    return cudaGetLastError();
}

} // namespace mjx
