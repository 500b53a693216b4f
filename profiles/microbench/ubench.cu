// ubench.cu -- issue-rate microbenchmarks on B200 (sm_100a) for the instructions K2's design
// depends on: scalar vs packed (f32x2) fp32 math, int<->float conversions, PRMT, SHFL, LDS.128.
// Each thread runs ITER iterations of a 4x-unrolled body over 8 independent dependency chains
// (32 dependent-free ops per loop trip, so loop overhead is < 10 %); one CTA of 512 threads
// (4 warps per SM sub-partition) per SM.  Every result feeds the next op of its chain, so
// nothing can be hoisted.  Reported: warp-instructions per cycle per SMSP (op instructions only).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench ubench.cu ; run: ./ubench
#include <cstdio>
#include <cuda_runtime.h>

#define ITER 256

template <int OP>
__global__ void __launch_bounds__(512) bench(float *out, long long *cycles, float seed, int iseed) {
    float a[8];
    unsigned long long d[8];
    int q[8];
#pragma unroll
    for(int i = 0; i < 8; i++) {
        a[i] = seed + threadIdx.x * 0.001f + i;
        float2 t = make_float2(a[i], a[i] * 0.5f);
        d[i] = *reinterpret_cast<unsigned long long *>(&t);
        q[i] = threadIdx.x * 77 + i * 13 + iseed;
    }
    const float c1 = 1.0001f, c2 = 0.0003f;
    float2 cc = make_float2(c1, c1), dd = make_float2(c2, c2);
    unsigned long long C1 = *reinterpret_cast<unsigned long long *>(&cc), C2 = *reinterpret_cast<unsigned long long *>(&dd);
    __shared__ float4 sm[512];
    sm[threadIdx.x] = make_float4(seed, seed, seed, seed);
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for(int it = 0; it < ITER; it++) {
#pragma unroll
        for(int u = 0; u < 4; u++) {
#pragma unroll
            for(int i = 0; i < 8; i++) {
                if(OP == 0) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(c1), "f"(c2));
                if(OP == 1) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(d[i]) : "l"(C1), "l"(C2));
                if(OP == 2) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(d[i]) : "l"(C2));
                if(OP == 3) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(c2));
                if(OP == 4) { asm volatile("cvt.rn.f32.s32 %0, %1;" : "=f"(a[i]) : "r"(q[i])); asm volatile("cvt.rzi.s32.f32 %0, %1;" : "=r"(q[i]) : "f"(a[i])); }
                if(OP == 5) { asm volatile("cvt.rn.f32.s32 %0, %1;" : "=f"(a[i]) : "r"(q[i])); asm volatile("mov.b32 %0, %1;" : "=r"(q[i]) : "f"(a[i])); }
                if(OP == 6) { asm volatile("cvt.rzi.s32.f32 %0, %1;" : "=r"(q[i]) : "f"(a[i])); asm volatile("mov.b32 %0, %1;" : "=f"(a[i]) : "r"(q[i])); }
                if(OP == 7) { asm volatile("prmt.b32 %0, %0, %1, 0x9910;" : "+r"(q[i]) : "r"(iseed)); }
                if(OP == 8) { q[i] = __shfl_xor_sync(0xffffffffu, q[i], 1); }
                if(OP == 9) { float4 v = sm[(q[i]) & 511]; q[i] = __float_as_int(v.x); }
                if(OP == 10) { asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(d[i]) : "l"(C1), "l"(C2));
                               asm volatile("prmt.b32 %0, %0, %1, 0x9910;" : "+r"(q[i]) : "r"(iseed)); }
                if(OP == 11) { asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(c1), "f"(c2));
                               asm volatile("prmt.b32 %0, %0, %1, 0x9910;" : "+r"(q[i]) : "r"(iseed)); }
                if(OP == 12) { short s; asm volatile("cvt.rzi.s16.f32 %0, %1;" : "=h"(s) : "f"(a[i])); asm volatile("cvt.rn.f32.s16 %0, %1;" : "=f"(a[i]) : "h"(s)); }
                if(OP == 13) { asm volatile("cvt.rzi.s32.f32 %0, %1;" : "=r"(q[i]) : "f"(a[i]));
                               asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(a[i]) : "f"(__int_as_float(q[i])), "f"(c1), "f"(c2)); }
                if(OP == 14) { asm volatile("mad.lo.s32 %0, %0, %1, %1;" : "+r"(q[i]) : "r"(iseed)); }
                if(OP == 15) { asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(d[i]) : "l"(C1), "l"(C2));
                               asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(c1), "f"(c2)); }
                if(OP == 16) { asm volatile("lop3.b32 %0, %0, %1, %1, 0x96;" : "+r"(q[i]) : "r"(iseed)); }
                if(OP == 17) { asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(c1), "f"(c2));
                               asm volatile("lop3.b32 %0, %0, %1, %1, 0x96;" : "+r"(q[i]) : "r"(iseed)); }
                if(OP == 18) { asm volatile("cvt.rzi.f32.f32 %0, %0;" : "+f"(a[i])); }
                if(OP == 19) { asm volatile("cvt.rn.f32.s32 %0, %1;" : "=f"(a[i]) : "r"(q[i])); asm volatile("cvt.rzi.s32.f32 %0, %1;" : "=r"(q[i]) : "f"(a[i]));
                               asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(d[i]) : "l"(C1), "l"(C2)); }
                if(OP == 20) { asm volatile("add.s32 %0, %0, %1;" : "+r"(q[i]) : "r"(iseed)); }
                if(OP == 21) { asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(c1)); }
                if(OP == 22) { asm volatile("shr.s32 %0, %0, 16;" : "+r"(q[i])); asm volatile("add.s32 %0, %0, %1;" : "+r"(q[i]) : "r"(iseed)); }
            }
        }
    }
    long long t1 = clock64();
    float s = 0;
#pragma unroll
    for(int i = 0; i < 8; i++) {
        float2 t = *reinterpret_cast<float2 *>(&d[i]);
        s += a[i] + t.x + t.y + q[i];
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if(threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(const char *name, int per_op) {
    float *out;
    long long *cyc;
    cudaMalloc(&out, 148 * 512 * 4);
    cudaMalloc(&cyc, 148 * 8);
    bench<OP><<<148, 512>>>(out, cyc, 1.0f, 3);
    bench<OP><<<148, 512>>>(out, cyc, 1.0f, 3);
    cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0;
    for(int i = 0; i < 148; i++) avg += h[i];
    avg /= 148;
    double ipc = 4.0 * ITER * 32 * per_op / avg; // 4 warps per SMSP
    printf("%-46s %8.0f cycles  %.3f warp-instr/clk/SMSP\n", name, avg, ipc);
    cudaFree(out);
    cudaFree(cyc);
}

int main() {
    run<0>("FFMA", 1);
    run<3>("FADD", 1);
    run<21>("FMUL", 1);
    run<1>("FFMA2 (fma.rn.f32x2)", 1);
    run<2>("FADD2 (add.rn.f32x2)", 1);
    run<15>("FFMA2 + FFMA (2 instr)", 2);
    run<4>("I2F.S32 + F2I.S32 chain (2 instr)", 2);
    run<5>("I2F.S32 (+mov)", 1);
    run<6>("F2I.S32 (+mov)", 1);
    run<12>("F2I.S16 + I2F.S16 chain (2 instr)", 2);
    run<18>("FRND.TRUNC", 1);
    run<13>("F2I + FFMA (2 instr)", 2);
    run<19>("I2F + F2I + FFMA2 (3 instr)", 3);
    run<7>("PRMT", 1);
    run<16>("LOP3", 1);
    run<20>("IADD", 1);
    run<22>("SHR + IADD (2 instr)", 2);
    run<14>("IMAD", 1);
    run<10>("FFMA2 + PRMT (2 instr)", 2);
    run<11>("FFMA + PRMT (2 instr)", 2);
    run<17>("FFMA + LOP3 (2 instr)", 2);
    run<8>("SHFL.BFLY", 1);
    run<9>("LDS.128 dependent chain (latency)", 1);
    return 0;
}
