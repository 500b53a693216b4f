// dct_ubench2.cu -- does interleaving the ALU-heavy halves of a block (int16 -> fp32 conversion, requantisation) with
// its FMA-heavy halves (the column passes) in SOURCE ORDER make ptxas / the in-order issue overlap them?
//   order A (what k2_generic_kernel does): convert all 32 pairs | 4 inverse column passes | rows | blend | rows |
//                                          4 forward column passes | requantise all 32 pairs
//   order B: for each column pair j: convert its 8 pairs, inverse column pass j | rows | blend | rows |
//            for each column pair j: forward column pass j, requantise its 8 pairs
// Register-only (no memory in the loop), 12 warps per SM like the kernel.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I ../../libmodjpeg_b200/csrc -I ../../include -o dct_ubench2 dct_ubench2.cu
#include <cstdio>
#include <cuda_runtime.h>

#include "mjx_math.cuh"
using namespace mjx;

__device__ __forceinline__ F2 cvt(uint32_t w) { return f2((float)(int)(short)(w & 0xffffu), (float)((int)w >> 16)); }

template <int ORDER>
__global__ void __launch_bounds__(128, 3) k(float *out, long long *cycles, int iters, uint32_t seed) {
    F2       x[32], y[32];
    uint32_t w[32], o[32];
#pragma unroll
    for(int i = 0; i < 32; i++) w[i] = (threadIdx.x * 2654435761u + i * 40503u + seed) & 0x00ff00ffu;
    long long t0 = clock64();
#pragma unroll 1
    for(int it = 0; it < iters; it++) {
        if(ORDER == 0) {
#pragma unroll
            for(int i = 0; i < 32; i++) x[i] = fma2(cvt(w[i]), f2(-0.3f, -0.4f), f2(1.0f + i, 2.0f));
#pragma unroll
            for(int j = 0; j < 4; j++) idct8p_cols_to_rowpairs(x, y, j);
        }
        else {
#pragma unroll
            for(int j = 0; j < 4; j++) {
#pragma unroll
                for(int r = 0; r < 8; r++) x[4 * r + j] = fma2(cvt(w[4 * r + j]), f2(-0.3f, -0.4f), f2(1.0f + r, 2.0f + j));
                idct8p_cols_to_rowpairs(x, y, j);
            }
        }
#pragma unroll
        for(int i = 0; i < 4; i++) idct8p<1>(y + 8 * i);
#pragma unroll
        for(int i = 0; i < 32; i++) y[i] = mul2(y[i], f2(0.5f + 0.001f * i, 0.25f));
#pragma unroll
        for(int i = 0; i < 4; i++) fdct8p_rowpairs_to_cols(y, x, i);
        if(ORDER == 0) {
#pragma unroll
            for(int j = 0; j < 4; j++) fdct8p<4>(x + j);
#pragma unroll
            for(int i = 0; i < 32; i++) o[i] = requant_pair(x[i], f2(0.0025f, 0.0031f), cvt(w[i]), f2(5.0f, 7.0f), f2(0.2000001f, 0.1428572f));
        }
        else {
#pragma unroll
            for(int j = 0; j < 4; j++) {
                fdct8p<4>(x + j);
#pragma unroll
                for(int r = 0; r < 8; r++)
                    o[4 * r + j] = requant_pair(x[4 * r + j], f2(0.0025f, 0.0031f), cvt(w[4 * r + j]), f2(5.0f, 7.0f), f2(0.2000001f, 0.1428572f));
            }
        }
#pragma unroll
        for(int i = 0; i < 32; i++) w[i] = (o[i] + w[(i + 1) & 31]) & 0x00ff00ffu; // the next "image"
    }
    long long t1 = clock64();
    uint32_t s = 0;
#pragma unroll
    for(int i = 0; i < 32; i++) s += w[i];
    out[blockIdx.x * 128 + threadIdx.x] = (float)s;
    if(threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int ORDER>
void run(const char *name) {
    float *out; long long *cyc;
    const int grid = 148 * 3, iters = 2000;
    cudaMalloc(&out, (size_t)grid * 128 * 4); cudaMalloc(&cyc, grid * 8);
    k<ORDER><<<grid, 128>>>(out, cyc, iters, 1u);
    k<ORDER><<<grid, 128>>>(out, cyc, iters, 1u);
    cudaDeviceSynchronize();
    long long *h = new long long[grid];
    cudaMemcpy(h, cyc, grid * 8, cudaMemcpyDeviceToHost);
    double avg = 0; for(int i = 0; i < grid; i++) avg += h[i]; avg /= grid;
    printf("%-60s %8.0f cycles per block of one warp -> %7.1f cycles per 32 blocks per scheduler (3 warps each)\n", name, avg / iters, avg / iters / 3);
    cudaFree(out); cudaFree(cyc); delete[] h;
}

int main() {
    run<0>("order A: phases (convert | passes | requantise)");
    run<1>("order B: column-pair interleaved");
    return 0;
}
