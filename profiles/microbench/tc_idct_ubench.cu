// tc_idct_ubench.cu -- unit test + timing of the tensor-core inverse half of the G class:
//   [128 blocks x 64 int16 coefficients]  x  B(table)  ->  TMEM: 64 pixel values (IDCT2(I*q) / 32) and 64 exact I*q / 512
// A operand: the raw int16 rows are converted IN PLACE to fp16 (I / 512) with one LOP3 + one HFMA2 per pair (valid for
// baseline-range coefficients -1024..1023) inside the canonical K-major SWIZZLE_128B layout; B = 16 * (C (x) C) diag(q) split
// into two fp16 pieces (hi, lo) plus diag(q) in rows 64..127 of the hi piece.  kind::f16, M = 128, N = 128 / 64, K = 16.
// Checks the result against double precision, reports the error beside the fp32 AAN path's, and times the MMA batch.
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I libmodjpeg_b200/csrc -I include -o tc_idct_ubench tc_idct_ubench.cu
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "mjx_math.cuh"

#define CK(x)                                                                                    \
    do {                                                                                         \
        cudaError_t e_ = (x);                                                                    \
        if(e_ != cudaSuccess) {                                                                  \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);      \
            exit(2);                                                                             \
        }                                                                                        \
    } while(0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// instruction descriptor, kind::f16: D fp32 (bits 4-5 = 1), A/B fp16 (0), both K-major, N >> 3 at bit 17, M >> 4 at bit 24
static constexpr uint32_t kIdescN128 = (1u << 4) | ((128u >> 3) << 17) | ((128u >> 4) << 24);
static constexpr uint32_t kIdescN64 = (1u << 4) | ((64u >> 3) << 17) | ((128u >> 4) << 24);

// shared-memory matrix descriptor: K-major, SWIZZLE_128B, 8-row groups 1024 B apart
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFFu);
    d |= (uint64_t)((1024u >> 4) & 0x3FFFu) << 32; // stride byte offset
    d |= (uint64_t)1 << 46;                        // descriptor version (sm_100)
    d |= (uint64_t)2 << 61;                        // SWIZZLE_128B
    return d;
}

__device__ __forceinline__ void mma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d), "l"(adesc),
                 "l"(bdesc), "r"(idesc), "r"(acc)
                 : "memory");
}
__device__ __forceinline__ void mma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// bounded wait: returns false when the phase did not complete (a wrong descriptor must not hang the box)
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity) {
    for(int spin = 0; spin < (1 << 22); spin++) {
        uint32_t ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if(ok) return true;
    }
    return false;
}
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

static constexpr int kABytes = 128 * 128;     // 128 rows x 64 fp16
static constexpr int kBHiBytes = 128 * 128;   // N = 128 rows
static constexpr int kBLoBytes = 64 * 128;    // N = 64 rows
static constexpr int kSmem = kABytes + kBHiBytes + kBLoBytes + 1024; // + alignment slack

// raw: [ncta][128 blocks][64] int16; bhi/blo: pre-swizzled operand images; out: [ncta][128][128] float
// iters > 1: timing mode (the MMA batch + TMEM read repeated, cycles per batch written to cyc[cta])
__global__ void __launch_bounds__(128) tc_idct_kernel(const int16_t *raw, const uint4 *bhi, const uint4 *blo, float *out, int iters, long long *cyc, int *err) {
    extern __shared__ unsigned char smem_dyn[];
    __shared__ uint32_t       s_tmem;
    __shared__ __align__(8) unsigned long long s_bar;
    unsigned char *base = (unsigned char *)(((uintptr_t)smem_dyn + 1023) & ~(uintptr_t)1023);
    unsigned char *sA = base, *sBh = base + kABytes, *sBl = sBh + kBHiBytes;
    const int      t = threadIdx.x, warp = t >> 5;

    if(warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "n"(128) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if(t == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&s_bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // B operand images (already in the canonical layout relative to a 1024-byte aligned base)
    for(int i = t; i < kBHiBytes / 16; i += 128) reinterpret_cast<uint4 *>(sBh)[i] = bhi[i];
    for(int i = t; i < kBLoBytes / 16; i += 128) reinterpret_cast<uint4 *>(sBl)[i] = blo[i];
    // raw rows: thread t owns row t; chunk c of row r lives at (r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4)
    unsigned char *myrow = sA + (t >> 3) * 1024 + (t & 7) * 128;
    const uint4   *src = reinterpret_cast<const uint4 *>(raw + ((size_t)(blockIdx.x & 7) * 128 + t) * 64);
    uint32_t kx; // kept in a register so that (w & imm) ^ kx is ONE LOP3
    asm volatile("mov.u32 %0, 0x04000400;" : "=r"(kx));
#pragma unroll
    for(int c = 0; c < 8; c++) {
        uint4 w = src[c];
        // int16 pair -> fp16 pair I / 512: (I + 1024) as a 11-bit field IS the fp16 bit pattern of (I + 1024) * 2^-24
        uint32_t *p = &w.x;
#pragma unroll
        for(int k = 0; k < 4; k++) {
            uint32_t u;
            asm("lop3.b32 %0, %1, 0x07FF07FF, %2, 0x6A;" : "=r"(u) : "r"(p[k]), "r"(kx)); // (w & mask) ^ kx
            asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(p[k]) : "r"(u), "r"(0x78007800u), "r"(0xC000C000u));
        }
        *reinterpret_cast<uint4 *>(myrow + ((c ^ (t & 7)) << 4)) = w;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s_tmem;
    const uint32_t bar = smem_u32(&s_bar);

    long long t0 = 0, t1 = 0;
    bool      ok = true;
    float     v[32];
    float     keep = 0.f;
    for(int it = 0; it < iters && ok; it++) {
        if(it == 1) t0 = clock64();
        if(t == 0) {
            const uint64_t ad = make_desc(smem_u32(sA)), bh = make_desc(smem_u32(sBh)), bl = make_desc(smem_u32(sBl));
#pragma unroll
            for(int k = 0; k < 4; k++) mma_f16(tmem, ad + 2 * k, bh + 2 * k, kIdescN128, k > 0);
#pragma unroll
            for(int k = 0; k < 4; k++) mma_f16(tmem, ad + 2 * k, bl + 2 * k, kIdescN64, 1);
            mma_commit(bar);
        }
        ok = mbar_wait(bar, it & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if(!ok) break;
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
#pragma unroll
        for(int g = 0; g < 4; g++) {
            tmem_ld32(taddr + 32 * g, v);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if(it == 0) {
                float *o = out + ((size_t)(blockIdx.x & 7) * 128 + t) * 128 + 32 * g;
#pragma unroll
                for(int i = 0; i < 32; i++) o[i] = v[i];
            }
            else {
#pragma unroll
                for(int i = 0; i < 32; i++) keep += v[i];
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads(); // everybody has read the accumulators before the next batch overwrites them
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    t1 = clock64();
    if(!ok) atomicExch(err, 1);
    if(t == 0 && iters > 1) cyc[blockIdx.x] = (t1 - t0) / (iters - 1);
    if(keep == 1234.5f) out[0] = keep;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if(warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(128) : "memory");
}

// ---------------------------------------------------------------------------------------------
static double Cm[8][8];
static void   init_c() {
    for(int k = 0; k < 8; k++)
        for(int n = 0; n < 8; n++) Cm[k][n] = (k == 0 ? sqrt(0.125) : 0.5) * cos((2 * n + 1) * k * M_PI / 16.0);
}
// column j of the product = pixel (row 2i + h, col k) with j = 2 * (8 i + k) + h  (the Q pairing of k2_compose.cu)
static void col_to_pixel(int j, int &y, int &x) {
    const int h = j & 1, p = j >> 1, i = p >> 3, k = p & 7;
    y = 2 * i + h, x = k;
}
static size_t swz(int row, int k) { // byte offset of fp16 element (row, k) in the canonical K-major SWIZZLE_128B layout
    const int chunk = k >> 3;
    return (size_t)(row >> 3) * 1024 + (row & 7) * 128 + ((chunk ^ (row & 7)) << 4) + (k & 7) * 2;
}

int main(int argc, char **argv) {
    init_c();
    int dev = 0;
    CK(cudaSetDevice(dev));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, dev));
    printf("device: %s, %d SMs\n", prop.name, prop.multiProcessorCount);

    // quantisation table: libjpeg standard luminance at quality 85 (natural order)
    static const int std_luma[64] = {16, 11, 10, 16, 24, 40, 51, 61, 12, 12, 14, 19, 26, 58, 60, 55, 14, 13, 16, 24, 40, 57, 69, 56, 14, 17, 22, 29, 51, 87, 80, 62,
                                     18, 22, 37, 56, 68, 109, 103, 77, 24, 35, 55, 64, 81, 104, 113, 92, 49, 64, 78, 87, 103, 121, 120, 101, 72, 92, 95, 98, 112, 100, 103, 99};
    for(int tcase = 0; tcase < 3; tcase++) {
        int q[64];
        for(int i = 0; i < 64; i++) {
            if(tcase == 0) q[i] = (std_luma[i] * 30 + 50) / 100 < 1 ? 1 : (std_luma[i] * 30 + 50) / 100; // q85
            else if(tcase == 1) q[i] = 1 + (rand() % 255);
            else q[i] = 255;
        }
        // operand images
        std::vector<__half> bhi(128 * 64), blo(64 * 64);
        std::vector<unsigned char> bhi_img(kBHiBytes, 0), blo_img(kBLoBytes, 0);
        for(int j = 0; j < 64; j++) {
            int y, x;
            col_to_pixel(j, y, x);
            for(int c = 0; c < 64; c++) {
                const int    v = c >> 3, u = c & 7;
                const double m = 16.0 * q[c] * Cm[v][y] * Cm[u][x];
                const __half hi = __float2half_rn((float)m);
                const __half lo = __float2half_rn((float)(m - (double)__half2float(hi)));
                memcpy(&bhi_img[swz(j, c)], &hi, 2);
                memcpy(&blo_img[swz(j, c)], &lo, 2);
            }
        }
        for(int c = 0; c < 64; c++) {
            const __half qh = __float2half_rn((float)q[c]);
            memcpy(&bhi_img[swz(64 + c, c)], &qh, 2);
        }
        // coefficients: case 0 JPEG-like (quantised DCT of smooth blocks + noise), others uniform over the baseline range
        const int ncta = 8, nblk = ncta * 128;
        std::vector<int16_t> I((size_t)nblk * 64);
        for(int b = 0; b < nblk; b++) {
            if(tcase == 0) {
                double px[8][8];
                const double base = (rand() % 200) - 100, gx = ((rand() % 200) - 100) / 10.0, gy = ((rand() % 200) - 100) / 10.0;
                for(int y = 0; y < 8; y++)
                    for(int x = 0; x < 8; x++) {
                        double p = base + gx * (x - 3.5) + gy * (y - 3.5) + ((rand() % 2001) - 1000) / 1000.0 * 12.0;
                        px[y][x] = p < -128 ? -128 : (p > 127 ? 127 : p);
                    }
                for(int v = 0; v < 8; v++)
                    for(int u = 0; u < 8; u++) {
                        double s = 0;
                        for(int y = 0; y < 8; y++)
                            for(int x = 0; x < 8; x++) s += Cm[v][y] * Cm[u][x] * px[y][x];
                        I[(size_t)b * 64 + v * 8 + u] = (int16_t)lrint(s / q[v * 8 + u]);
                    }
            }
            else
                for(int c = 0; c < 64; c++) I[(size_t)b * 64 + c] = (int16_t)((rand() % 2048) - 1024);
        }
        int16_t   *d_raw;
        uint4     *d_bhi, *d_blo;
        float     *d_out;
        long long *d_cyc;
        int       *d_err;
        CK(cudaMalloc(&d_raw, I.size() * 2));
        CK(cudaMalloc(&d_bhi, kBHiBytes));
        CK(cudaMalloc(&d_blo, kBLoBytes));
        CK(cudaMalloc(&d_out, (size_t)nblk * 128 * 4));
        CK(cudaMalloc(&d_cyc, 2048 * 8));
        CK(cudaMalloc(&d_err, 4));
        CK(cudaMemset(d_err, 0, 4));
        CK(cudaMemcpy(d_raw, I.data(), I.size() * 2, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(d_bhi, bhi_img.data(), kBHiBytes, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(d_blo, blo_img.data(), kBLoBytes, cudaMemcpyHostToDevice));
        CK(cudaFuncSetAttribute(tc_idct_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
        tc_idct_kernel<<<ncta, 128, kSmem>>>(d_raw, d_bhi, d_blo, d_out, 1, d_cyc, d_err);
        CK(cudaDeviceSynchronize());
        int herr = 0;
        CK(cudaMemcpy(&herr, d_err, 4, cudaMemcpyDeviceToHost));
        if(herr) {
            printf("case %d: MMA did not complete (mbarrier timeout)\n", tcase);
            return 3;
        }
        std::vector<float> out((size_t)nblk * 128);
        CK(cudaMemcpy(out.data(), d_out, out.size() * 4, cudaMemcpyDeviceToHost));

        // reference in double; the fp32 AAN path of k2 (dequantise with prescale, two idct8s passes) beside it
        double emax_tc = 0, esum_tc = 0, emax_aan = 0, esum_aan = 0, vmax = 0;
        long   iq_bad = 0;
        for(int b = 0; b < nblk; b++) {
            double ref[8][8];
            float  xa[64];
            for(int c = 0; c < 64; c++) xa[c] = (float)I[(size_t)b * 64 + c] * ((float)q[c] * mjx::inv_scale(c >> 3) * mjx::inv_scale(c & 7));
            for(int u = 0; u < 8; u++) mjx::idct8s<8>(xa + u);
            for(int v = 0; v < 8; v++) mjx::idct8s<1>(xa + 8 * v);
            for(int y = 0; y < 8; y++)
                for(int x = 0; x < 8; x++) {
                    double s = 0;
                    for(int v = 0; v < 8; v++)
                        for(int u = 0; u < 8; u++) s += Cm[v][y] * Cm[u][x] * (double)I[(size_t)b * 64 + v * 8 + u] * q[v * 8 + u];
                    ref[y][x] = s;
                }
            for(int j = 0; j < 64; j++) {
                int y, x;
                col_to_pixel(j, y, x);
                const double got = (double)out[(size_t)b * 128 + j] * 32.0;
                const double e = fabs(got - ref[y][x]), ea = fabs((double)xa[y * 8 + x] - ref[y][x]);
                emax_tc = fmax(emax_tc, e), esum_tc += e * e;
                emax_aan = fmax(emax_aan, ea), esum_aan += ea * ea;
                vmax = fmax(vmax, fabs(ref[y][x]));
            }
            for(int c = 0; c < 64; c++)
                if(out[(size_t)b * 128 + 64 + c] * 512.0f != (float)((int)I[(size_t)b * 64 + c] * q[c])) iq_bad++;
        }
        printf("case %d: |pixel| max %.1f   tensor-core: max err %.3e rms %.3e   fp32 AAN: max err %.3e rms %.3e   I*q mismatches %ld of %d\n", tcase, vmax, emax_tc,
               sqrt(esum_tc / (nblk * 64.0)), emax_aan, sqrt(esum_aan / (nblk * 64.0)), iq_bad, nblk * 64);

        if(tcase == 0) {
            // timing: one CTA alone, then 3 CTAs on every SM
            for(int grid : {1, prop.multiProcessorCount, 3 * prop.multiProcessorCount}) {
                CK(cudaMemset(d_cyc, 0, 2048 * 8));
                tc_idct_kernel<<<grid, 128, kSmem>>>(d_raw, d_bhi, d_blo, d_out, 2001, d_cyc, d_err);
                CK(cudaDeviceSynchronize());
                std::vector<long long> cyc(grid);
                CK(cudaMemcpy(cyc.data(), d_cyc, grid * 8, cudaMemcpyDeviceToHost));
                long long mn = cyc[0], mx = cyc[0];
                for(long long c : cyc) mn = c < mn ? c : mn, mx = c > mx ? c : mx;
                printf("  timing grid %d (every CTA converts the same 128 rows): cycles per batch (8 MMAs + commit + wait + 128-column TMEM read) min %lld max %lld\n", grid, mn, mx);
            }
        }
        cudaFree(d_raw), cudaFree(d_bhi), cudaFree(d_blo), cudaFree(d_out), cudaFree(d_cyc), cudaFree(d_err);
    }
    return 0;
}
