// dct_ubench.cu -- how fast can the packed-fp32 block arithmetic of k2_generic_kernel run when NOTHING else is in the
// way?  Each thread keeps one block in registers and repeats IDCT (two pairings) -> multiply -> FDCT on it, no memory
// traffic in the loop.  Reported per configuration (warps per SM): FMA-pipe instructions per clock per SM sub-partition.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I ../../libmodjpeg_b200/csrc -I ../../include -o dct_ubench dct_ubench.cu
#include <cstdio>
#include <cuda_runtime.h>

#include "mjx_math.cuh"
using namespace mjx;

template <int THREADS, int MINB, bool WITH_ALU>
__global__ void __launch_bounds__(THREADS, MINB) k(float *out, long long *cycles, int iters, float seed) {
    F2 x[32], y[32], a[32];
    uint32_t wsrc = threadIdx.x * 2654435761u + 12345u;
#pragma unroll
    for(int i = 0; i < 32; i++) {
        x[i] = f2(seed + i + threadIdx.x * 0.01f, seed - i);
        a[i] = f2(0.5f + 0.001f * i, 0.25f);
    }
    long long t0 = clock64();
#pragma unroll 1
    for(int it = 0; it < iters; it++) {
#pragma unroll
        for(int j = 0; j < 4; j++) idct8p_cols_to_rowpairs(x, y, j);
#pragma unroll
        for(int i = 0; i < 4; i++) idct8p<1>(y + 8 * i);
#pragma unroll
        for(int i = 0; i < 32; i++) y[i] = mul2(y[i], a[i]);
#pragma unroll
        for(int i = 0; i < 4; i++) fdct8p_rowpairs_to_cols(y, x, i);
#pragma unroll
        for(int j = 0; j < 4; j++) fdct8p<4>(x + j);
#pragma unroll
        for(int i = 0; i < 32; i++) x[i] = mul2(x[i], bc2(0.015625f)); // keep the values bounded
        if(WITH_ALU) { // the ALU-heavy part of the block: int16 -> fp32 conversions and the requantisation of every pair
#pragma unroll
            for(int i = 0; i < 32; i++) {
                const uint32_t w = wsrc ^ (uint32_t)(i * 0x01010101u);
                const F2       I = f2((float)(int)(short)(w & 0xffffu), (float)((int)w >> 16));
                const uint32_t o = requant_pair(x[i], f2(0.25f, 0.31f), I, f2(5.0f, 7.0f), f2(0.2000001f, 0.1428572f));
                x[i] = fma2(I, f2(-0.3f, -0.4f), f2((float)(int)(short)(o & 0xffffu) * 1e-3f, (float)((int)o >> 16) * 1e-3f));
                wsrc = wsrc * 1664525u + o;
            }
        }
    }
    long long t1 = clock64();
    float s = 0;
#pragma unroll
    for(int i = 0; i < 32; i++) s += x[i].x + x[i].y;
    s += (float)(wsrc & 0xff);
    out[blockIdx.x * THREADS + threadIdx.x] = s;
    if(threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int THREADS, int MINB, bool WITH_ALU>
void run(int ctas_per_sm) {
    float *out; long long *cyc;
    const int grid = 148 * ctas_per_sm, iters = 2000;
    cudaMalloc(&out, (size_t)grid * THREADS * 4); cudaMalloc(&cyc, grid * 8);
    k<THREADS, MINB, WITH_ALU><<<grid, THREADS>>>(out, cyc, iters, 1.0f);
    k<THREADS, MINB, WITH_ALU><<<grid, THREADS>>>(out, cyc, iters, 1.0f);
    cudaDeviceSynchronize();
    long long *h = new long long[grid];
    cudaMemcpy(h, cyc, grid * 8, cudaMemcpyDeviceToHost);
    double avg = 0; for(int i = 0; i < grid; i++) avg += h[i]; avg /= grid;
    // per iteration and thread: 16 x 30 packed-or-scalar FMA-pipe instructions + 64 scalar transition extras = 544 + 32 + 32 mul2
    const double fma_instr = 544.0 + 64.0;
    const double warps_per_smsp = (double)THREADS / 32 * ctas_per_sm / 4;
    printf("%s %2d warps/SM (%d x %d threads): %8.0f cycles per block-iteration of one warp -> %7.1f cycles per warp-block per SMSP, %.3f DCT FMA-pipe instr/clk/SMSP\n",
           WITH_ALU ? "DCT + convert + requant" : "DCT only              ", THREADS / 32 * ctas_per_sm, ctas_per_sm, THREADS, avg / iters, avg / iters / warps_per_smsp,
           fma_instr * warps_per_smsp / (avg / iters));
    cudaFree(out); cudaFree(cyc); delete[] h;
}

int main() {
    run<128, 2, false>(2);
    run<128, 3, false>(3);
    run<128, 4, false>(4);
    run<128, 2, true>(2);
    run<128, 3, true>(3);
    run<128, 4, true>(4);
    run<128, 5, true>(5);
    run<128, 6, true>(6);
    return 0;
}
