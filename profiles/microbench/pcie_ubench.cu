// pcie_ubench.cu -- what the zero-copy path of mjx_compose_batch_host can expect from the PCIe link:
// copy-engine bandwidth (one direction, both directions) against kernels that read / write page-locked
// host memory directly in 128-byte blocks (one JPEG coefficient block), dense or with the sparse pattern
// of a tiled logo (62 % of the blocks written, 40 % read), with different store flavours.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pcie_ubench pcie_ubench.cu ; run: ./pcie_ubench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if(e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while(0)

__device__ __forceinline__ bool keep(unsigned b, unsigned pct) { return ((b * 2654435761u) >> 16) % 100u < pct; }

// MODE 0: st.cs 16 B per lane (8 lanes per block)   1: default st   2: bulk (TMA) store, one lane per block
// 3: ld 16 B per lane    4: ld + st (read-modify-write)
template <int MODE>
__global__ void __launch_bounds__(256) k(uint4 *host, size_t nblocks, unsigned pct, unsigned long long *sink) {
    __shared__ __align__(128) uint4 stage[8][32 * 8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t nw = (size_t)gridDim.x * 8, w = (size_t)blockIdx.x * 8 + warp;
    unsigned long long acc = 0;
    for(size_t base = w * 32; base < nblocks; base += nw * 32) {
        if(MODE == 2) {
            for(int i = 0; i < 8; i++) stage[warp][lane * 8 + i] = make_uint4(lane, i, (unsigned)base, 7);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            const size_t b = base + lane;
            if(b < nblocks && keep((unsigned)b, pct)) {
                unsigned src = (unsigned)__cvta_generic_to_shared(&stage[warp][lane * 8]);
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], 128;" ::"l"(host + b * 8), "r"(src) : "memory");
            }
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            __syncwarp();
            continue;
        }
#pragma unroll
        for(int j = 0; j < 8; j++) {
            const size_t b = base + j * 4 + (lane >> 3);
            if(b >= nblocks || !keep((unsigned)b, pct)) continue;
            uint4 *p = host + b * 8 + (lane & 7);
            if(MODE == 0) __stcs(p, make_uint4(lane, j, (unsigned)b, 1));
            if(MODE == 1) *p = make_uint4(lane, j, (unsigned)b, 1);
            if(MODE == 3) { uint4 v = __ldcs(p); acc += v.x + v.w; }
            if(MODE == 4) { uint4 v = __ldcs(p); v.x += 1; __stcs(p, v); }
        }
    }
    if(MODE == 3 && acc == 0x123456789ull) *sink = acc;
}

template <int MODE>
int run(const char *name, uint4 *host, size_t nblocks, unsigned pct, unsigned long long *sink, int rw) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    k<MODE><<<148 * 4, 256>>>(host, nblocks, pct, sink);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    k<MODE><<<148 * 4, 256>>>(host, nblocks, pct, sink);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    size_t cnt = 0;
    for(size_t b = 0; b < nblocks; b++) cnt += (((unsigned)b * 2654435761u) >> 16) % 100u < pct;
    printf("%-44s pct %3u  %8.3f ms  %7.2f GB/s per direction%s\n", name, pct, ms, cnt * 128.0 / ms / 1e6, rw ? " (read AND write)" : "");
    return 0;
}

int main() {
    const size_t bytes = 1ull << 30, nblocks = bytes / 128;
    uint4 *host; void *dev, *dev2; uint4 *host2; unsigned long long *sink;
    CK(cudaHostAlloc(&host, bytes, cudaHostAllocPortable));
    CK(cudaHostAlloc(&host2, bytes, cudaHostAllocPortable));
    CK(cudaMalloc(&dev, bytes)); CK(cudaMalloc(&dev2, bytes)); CK(cudaMalloc(&sink, 8));
    for(size_t i = 0; i < bytes / 16; i += 256) host[i] = make_uint4(1, 2, 3, 4), host2[i] = make_uint4(1, 2, 3, 4);
    cudaStream_t s1, s2; CK(cudaStreamCreate(&s1)); CK(cudaStreamCreate(&s2));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float ms;
    for(int rep = 0; rep < 2; rep++) {
        CK(cudaEventRecord(e0, s1)); CK(cudaMemcpyAsync(dev, host, bytes, cudaMemcpyHostToDevice, s1)); CK(cudaEventRecord(e1, s1)); CK(cudaStreamSynchronize(s1));
        CK(cudaEventElapsedTime(&ms, e0, e1)); if(rep) printf("copy engine H2D 1 GiB                        %8.3f ms  %7.2f GB/s\n", ms, bytes / ms / 1e6);
        CK(cudaEventRecord(e0, s1)); CK(cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, s1)); CK(cudaEventRecord(e1, s1)); CK(cudaStreamSynchronize(s1));
        CK(cudaEventElapsedTime(&ms, e0, e1)); if(rep) printf("copy engine D2H 1 GiB                        %8.3f ms  %7.2f GB/s\n", ms, bytes / ms / 1e6);
        CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(e0, s1));
        CK(cudaMemcpyAsync(dev, host, bytes, cudaMemcpyHostToDevice, s1));
        CK(cudaMemcpyAsync(host2, dev2, bytes, cudaMemcpyDeviceToHost, s2));
        CK(cudaStreamSynchronize(s2));
        CK(cudaEventRecord(e1, s1)); CK(cudaStreamSynchronize(s1));
        CK(cudaEventElapsedTime(&ms, e0, e1)); if(rep) printf("copy engine H2D + D2H concurrently           %8.3f ms  %7.2f GB/s per direction\n", ms, bytes / ms / 1e6);
    }
    for(unsigned pct : {100u, 62u}) {
        if(run<0>("kernel st.cs 16 B/lane, 128 B blocks", host, nblocks, pct, sink, 0)) return 1;
        if(run<1>("kernel st (default) 16 B/lane", host, nblocks, pct, sink, 0)) return 1;
        if(run<2>("kernel bulk (TMA) store 128 B/lane", host, nblocks, pct, sink, 0)) return 1;
    }
    for(unsigned pct : {100u, 40u}) {
        if(run<3>("kernel ld.cs 16 B/lane, 128 B blocks", host, nblocks, pct, sink, 0)) return 1;
        if(run<4>("kernel ld.cs + st.cs (rmw)", host, nblocks, pct, sink, 1)) return 1;
    }
    return 0;
}
