// pipe_ubench.cu -- the data-movement skeleton of k2_generic_op_kernel without any arithmetic: per warp a ring of `stages` 4 KB
// shared-memory buffers filled by cp.async (32 rows of 128 bytes: both blocks of a slot pair for 16 images), drained by
// LDS.128 + coalesced STG.128, requests `depth` units ahead.  Optional coupling: the four warps of a group meet at a named
// barrier before a unit is drained (what the UMMA hand-over does to them).  What does the memory system deliver for this shape?
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_ubench pipe_ubench.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if(e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while(0)

template <bool CG>
__device__ __forceinline__ void cp16(unsigned dst, const void *src) {
    if(CG) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
    else asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// units of a CTA: pair p = blockIdx.x + j * gridDim.x (j = 0, 1, ..), batch b of 64 images; unit (j, b) -> group (b % ngroups);
// inside a group warp wq takes images 16 wq .. 16 wq + 15
template <int STAGES, bool COUPLE, bool CG = false, int ST = 0>
__global__ void __launch_bounds__(1024) k_pipe(char *base, size_t image_bytes, const unsigned *list, int S, int N) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, ngroups = blockDim.x >> 7, grp = warp >> 2, wq = warp & 3;
    const unsigned sbase = (unsigned)__cvta_generic_to_shared(smem) + warp * STAGES * 4096;
    const int npairs = S / 2, nb = (N + 63) / 64;
    const int bpg = grp < nb ? (nb - grp + ngroups - 1) / ngroups : 0;
    if(bpg == 0) return;
    const int my_pairs = blockIdx.x < npairs ? (npairs - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const int units = my_pairs * bpg;
    // lane -> 16-byte chunk (lane & 7) of row ((lane >> 3) & 1 = slot of the pair, lane >> 4 = image of the instruction's two)
    const int g_half = (lane >> 3) & 1, g_img = lane >> 4;
    auto addr_of = [&](int u, int i) -> char * {
        const int j = u / bpg, k = u - j * bpg;
        const int pr = blockIdx.x + j * gridDim.x;
        int       img = (grp + k * ngroups) * 64 + 16 * wq + 2 * i + g_img;
        if(img >= N) img = N - 1;
        return base + (size_t)img * image_bytes + (size_t)list[2 * pr + g_half] * 128 + (lane & 7) * 16;
    };
    auto request = [&](int u, int st) {
        if(u < units) {
#pragma unroll
            for(int i = 0; i < 8; i++) cp16<CG>(sbase + st * 4096 + i * 512 + lane * 16, addr_of(u, i));
        }
        cp_commit();
    };
#pragma unroll
    for(int d = 0; d < STAGES - 1; d++) request(d, d);
    for(int u = 0; u < units; u++) {
        const int st = u % STAGES;
        request(u + STAGES - 1, (u + STAGES - 1) % STAGES);
        cp_wait<STAGES - 1>();
        __syncwarp();
        if(COUPLE) asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");
        uint4 v[8];
#pragma unroll
        for(int i = 0; i < 8; i++) asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v[i].x), "=r"(v[i].y), "=r"(v[i].z), "=r"(v[i].w) : "r"(sbase + st * 4096 + i * 512 + lane * 16));
#pragma unroll
        for(int i = 0; i < 8; i++) {
            v[i].x += 1;
            if(ST == 0) __stcs((uint4 *)addr_of(u, i), v[i]);
            else if(ST == 1) __stcg((uint4 *)addr_of(u, i), v[i]);
            else __stwt((uint4 *)addr_of(u, i), v[i]);
        }
        __syncwarp();
    }
    cp_wait<0>();
}

int main() {
    const int    N = 1250;
    const size_t image_bytes = 6266880;
    const int    blocks_per_image = (int)(image_bytes / 128);
    std::vector<unsigned> list;
    for(int b = 0; b < blocks_per_image; b++)
        if(b % 30 < 12) list.push_back(b);
    int S = (int)list.size() / 32 * 32;
    list.resize(S);
    char *base;
    CK(cudaMalloc(&base, image_bytes * N));
    CK(cudaMemset(base, 1, image_bytes * N));
    unsigned *dl;
    CK(cudaMalloc(&dl, S * 4));
    CK(cudaMemcpy(dl, list.data(), S * 4, cudaMemcpyHostToDevice));
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0), cudaEventCreate(&e1);
    const double bytes = (double)S * N * 256.0;
    printf("%d images x %d listed blocks = %.2f GB read + written per pass\n", N, S, bytes / 1e9);
    auto run = [&](auto kernel, const char *name, int stages, int warps, int smem_force = 0) {
        const int smem = smem_force ? smem_force : warps * stages * 4096;
        CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        float best = 1e9f;
        for(int rep = 0; rep < 5; rep++) {
            cudaEventRecord(e0);
            kernel<<<sms, warps * 32, smem>>>(base, image_bytes, dl, S, N);
            cudaEventRecord(e1);
            CK(cudaEventSynchronize(e1));
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            if(rep > 0 && ms < best) best = ms;
        }
        printf("%-44s stages %d warps %2d smem %3d KB (%3d KB in flight per SM): %.3f ms  %.0f GB/s\n", name, stages, warps, smem / 1024, warps * (stages - 1) * 4, best, bytes / best / 1e6);
    };
    run(k_pipe<3, true>, "coupled, .ca loads, st.cs", 3, 12);
    // the same with the shared-memory carve-out of k2_generic_op_kernel: what is left of the 256 KB is the L1
    for(int kb : {160, 200, 227}) {
        run(k_pipe<3, true>, "coupled, .ca loads, st.cs", 3, 12, kb * 1024);
        run(k_pipe<3, true, true>, "coupled, .cg loads, st.cs", 3, 12, kb * 1024);
    }
    run(k_pipe<3, true, true, 1>, "coupled, .cg loads, st.cg", 3, 12, 227 * 1024);
    run(k_pipe<3, true, true, 2>, "coupled, .cg loads, st.wt", 3, 12, 227 * 1024);
    run(k_pipe<3, false, true>, "independent, .cg loads, st.cs", 3, 12, 227 * 1024);
    run(k_pipe<3, false, true>, "independent, .cg loads, st.cs", 3, 16, 227 * 1024);
    return 0;
}
