// tma_ubench.cu -- can per-lane 128-byte bulk (TMA) copies replace LDGSTS + LDS/STG for streaming 8x8
// coefficient blocks through shared memory at HBM speed?  Each warp moves 32 blocks (4 KB) per iteration:
// global -> shared -> global (a copy through smem, like K2's G kernel without the math).
//   mode 0: cp.async (LDGSTS 16 B/lane x 8) in, LDS.128 + STG.128 out   (what k2_generic_kernel does)
//   mode 1: one cp.async.bulk 128 B per lane in (mbarrier), one cp.async.bulk 128 B per lane out
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_ubench tma_ubench.cu ; run: ./tma_ubench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if(e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while(0)
static constexpr int kStride = 144;

__device__ __forceinline__ unsigned s32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

template <int MODE, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) k(const uint4 *in, uint4 *out, size_t nblocks) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) unsigned long long mbar[WARPS][2];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char *ws = smem + warp * 2 * 32 * kStride;
    const size_t nw = (size_t)gridDim.x * WARPS, w = (size_t)blockIdx.x * WARPS + warp;
    if(MODE == 1) {
        if(lane == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&mbar[warp][0])));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&mbar[warp][1])));
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        __syncwarp();
    }
    int it = 0;
    auto load = [&](size_t base, int st) {
        if(MODE == 0) {
#pragma unroll
            for(int j = 0; j < 8; j++) {
                const int t = j * 4 + (lane >> 3), c = lane & 7;
                asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(s32(ws + st * 32 * kStride + t * kStride + c * 16)), "l"(in + (base + t) * 8 + c) : "memory");
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        }
        else {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if(lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&mbar[warp][st])), "r"(32 * 128) : "memory");
            __syncwarp();
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 128, [%2];" ::"r"(s32(ws + st * 32 * kStride + lane * kStride)),
                         "l"(in + (base + lane) * 8), "r"(s32(&mbar[warp][st])) : "memory");
        }
    };
    size_t base = w * 32;
    if(base + 32 <= nblocks) load(base, 0);
    for(; base + 32 <= nblocks; base += nw * 32, it++) {
        const int st = it & 1;
        const size_t next = base + nw * 32;
        if(MODE == 1) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); // stage st^1 was the source of the previous store
        if(next + 32 <= nblocks) load(next, st ^ 1);
        else if(MODE == 0) asm volatile("cp.async.commit_group;" ::: "memory");
        if(MODE == 0) {
            asm volatile("cp.async.wait_group 1;" ::: "memory");
            __syncwarp();
        }
        else {
            const unsigned parity = (it >> 1) & 1;
            unsigned done = 0;
            while(!done) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(s32(&mbar[warp][st])), "r"(parity) : "memory");
        }
        // touch the block like the real kernel does (thread-per-block read + write back in place)
        uint4 *mine = reinterpret_cast<uint4 *>(ws + st * 32 * kStride + lane * kStride);
#pragma unroll
        for(int r = 0; r < 8; r++) { uint4 v = mine[r]; v.x ^= 1u; mine[r] = v; }
        if(MODE == 0) {
            __syncwarp();
#pragma unroll
            for(int j = 0; j < 8; j++) {
                const int t = j * 4 + (lane >> 3), c = lane & 7;
                const uint4 v = *reinterpret_cast<const uint4 *>(ws + st * 32 * kStride + t * kStride + c * 16);
                __stcs(out + (base + t) * 8 + c, v);
            }
            __syncwarp();
        }
        else {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], 128;" ::"l"(out + (base + lane) * 8), "r"(s32(mine)) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
    }
    if(MODE == 1) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

template <int MODE, int WARPS>
int run(const char *name, const uint4 *in, uint4 *out, size_t nblocks, int ctas_per_sm) {
    const int smem = WARPS * 2 * 32 * kStride;
    CK(cudaFuncSetAttribute(k<MODE, WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    k<MODE, WARPS><<<148 * ctas_per_sm, WARPS * 32, smem>>>(in, out, nblocks);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    k<MODE, WARPS><<<148 * ctas_per_sm, WARPS * 32, smem>>>(in, out, nblocks);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("%-52s %d warps/SM  %8.3f ms  %7.1f GB/s (read + write)\n", name, WARPS * ctas_per_sm, ms, 2.0 * nblocks * 128 / ms / 1e6);
    return 0;
}

int main() {
    const size_t bytes = 4ull << 30, nblocks = bytes / 128;
    uint4 *in, *out;
    CK(cudaMalloc(&in, bytes)); CK(cudaMalloc(&out, bytes));
    CK(cudaMemset(in, 1, bytes));
    if(run<0, 4>("LDGSTS in, LDS+STG out", in, out, nblocks, 3)) return 1;
    if(run<1, 4>("bulk (TMA) 128 B per lane in and out", in, out, nblocks, 3)) return 1;
    if(run<0, 4>("LDGSTS in, LDS+STG out", in, out, nblocks, 6)) return 1;
    if(run<1, 4>("bulk (TMA) 128 B per lane in and out", in, out, nblocks, 6)) return 1;
    // verify mode 1 copied correctly
    uint4 h[8];
    CK(cudaMemcpy(h, out + 12345 * 8, sizeof(h), cudaMemcpyDeviceToHost));
    printf("check: %08x %08x (expect 01010100 01010101)\n", h[0].x, h[0].y);
    return 0;
}
