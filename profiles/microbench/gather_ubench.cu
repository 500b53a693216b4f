// gather_ubench.cu -- what the memory system delivers for K2's two access orders, without any arithmetic.
// A batch of N images (6.27 MB apart), S listed 128-byte blocks per image (every block of a run pattern like the bench's
// G class).  Every listed block of every image is read and written back once (256 B of traffic per block), in one of two orders:
//   order 0 ("tile"):     a warp takes 32 consecutive list entries of ONE image (what k2_generic_kernel does)
//   order 1 ("operator"): a warp takes ONE list entry of 32 consecutive images (what k2_generic_op_kernel does);
//                         CTA c walks list entries c, c + grid, ...; inside an entry the images go in batches of 128 per 4 warps
//   order 2 ("operator, slot-window"): like 1, but the CTAs sweep a window of W list entries for one batch of 128 images
//                         before moving to the next batch (the same entries come back for every batch)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_ubench gather_ubench.cu ; run: ./gather_ubench
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if(e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while(0)

__device__ __forceinline__ void copy8(const uint4 *const (&src)[8], uint4 *const (&dst)[8]) {
    uint4 v[8];
#pragma unroll
    for(int i = 0; i < 8; i++) v[i] = __ldcs(src[i]);
#pragma unroll
    for(int i = 0; i < 8; i++) {
        v[i].x += 1;
        __stcs(dst[i], v[i]);
    }
}

// lane -> 16-byte chunk (lane & 7) of rows (lane >> 3) + 4 i
__global__ void __launch_bounds__(512) k_tile(char *base, size_t image_bytes, const unsigned *list, int S, int N) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const long long ntiles = (long long)(S / 32);
    const long long items = ntiles * N;
    for(long long it = (long long)blockIdx.x * nw + warp; it < items; it += (long long)gridDim.x * nw) {
        const long long tile = it / N;
        const int       img = (int)(it - tile * N);
        const uint4 *src[8];
        uint4       *dst[8];
#pragma unroll
        for(int i = 0; i < 8; i++) {
            const unsigned off = list[tile * 32 + (lane >> 3) + 4 * i];
            char *p = base + (size_t)img * image_bytes + (size_t)off * 128 + (lane & 7) * 16;
            src[i] = (const uint4 *)p, dst[i] = (uint4 *)p;
        }
        copy8(src, dst);
    }
}

__global__ void __launch_bounds__(512) k_op(char *base, size_t image_bytes, const unsigned *list, int S, int N, int window) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int nb32 = (N + 31) / 32; // warp-batches of 32 images
    if(window <= 0) {
        for(int s = blockIdx.x; s < S; s += gridDim.x) {
            const unsigned off = list[s];
            for(int b = warp; b < nb32; b += nw) {
                const uint4 *src[8];
                uint4       *dst[8];
#pragma unroll
                for(int i = 0; i < 8; i++) {
                    int img = b * 32 + (lane >> 3) + 4 * i;
                    if(img >= N) img = N - 1;
                    char *p = base + (size_t)img * image_bytes + (size_t)off * 128 + (lane & 7) * 16;
                    src[i] = (const uint4 *)p, dst[i] = (uint4 *)p;
                }
                copy8(src, dst);
            }
        }
    }
    else {
        // windows of `window` entries; inside a window: image batches of 32 * nw images outermost, entries dealt to the CTAs
        const int per = 32 * nw;
        for(int w0 = 0; w0 < S; w0 += window) {
            const int w1 = min(S, w0 + window);
            for(int b0 = 0; b0 < N; b0 += per) {
                for(int s = w0 + blockIdx.x; s < w1; s += gridDim.x) {
                    const unsigned off = list[s];
                    const uint4 *src[8];
                    uint4       *dst[8];
#pragma unroll
                    for(int i = 0; i < 8; i++) {
                        int img = b0 + warp * 32 + (lane >> 3) + 4 * i;
                        if(img >= N) img = N - 1;
                        char *p = base + (size_t)img * image_bytes + (size_t)off * 128 + (lane & 7) * 16;
                        src[i] = (const uint4 *)p, dst[i] = (uint4 *)p;
                    }
                    copy8(src, dst);
                }
            }
        }
    }
}

// order 3 ("operator, runs"): CTA c owns runs of R adjacent list entries (run index c, c + grid, ...); inside a run the units
// (entry r of the run, batch b of 128 images) go batch-major -- u = b * R + r -- and unit u to warp group u mod (nw / 4): what a CTA
// has in flight at any time are adjacent blocks of the same images
__global__ void __launch_bounds__(512) k_runs(char *base, size_t image_bytes, const unsigned *list, int S, int N, int R) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, ng = blockDim.x >> 7, grp = warp >> 2, wq = warp & 3;
    const int nb = (N + 127) / 128;
    for(int s0 = blockIdx.x * R; s0 < S; s0 += gridDim.x * R) {
        const int r_n = min(R, S - s0);
        for(int u = grp; u < nb * r_n; u += ng) {
            const int b = u / r_n, r = u - b * r_n;
            const unsigned off = list[s0 + r];
            const uint4 *src[8];
            uint4       *dst[8];
#pragma unroll
            for(int i = 0; i < 8; i++) {
                int img = b * 128 + wq * 32 + (lane >> 3) + 4 * i;
                if(img >= N) img = N - 1;
                char *p = base + (size_t)img * image_bytes + (size_t)off * 128 + (lane & 7) * 16;
                src[i] = (const uint4 *)p, dst[i] = (uint4 *)p;
            }
            copy8(src, dst);
        }
    }
}

// order 4 ("operator, R adjacent entries at once"): a warp's 32 rows = R adjacent list entries x 32 / R consecutive images, the R
// entries of an image in consecutive rows (so that one warp instruction asks for them together)
__global__ void __launch_bounds__(512) k_adj(char *base, size_t image_bytes, const unsigned *list, int S, int N, int R) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int ipw = 32 / R; // images per warp
    const int nbw = (N + ipw - 1) / ipw;
    for(int s0 = blockIdx.x * R; s0 + R <= S; s0 += gridDim.x * R) {
        for(int b = warp; b < nbw; b += nw) {
            const uint4 *src[8];
            uint4       *dst[8];
#pragma unroll
            for(int i = 0; i < 8; i++) {
                const int row = (lane >> 3) + 4 * i;
                int       img = b * ipw + row / R;
                if(img >= N) img = N - 1;
                const unsigned off = list[s0 + row % R];
                char *p = base + (size_t)img * image_bytes + (size_t)off * 128 + (lane & 7) * 16;
                src[i] = (const uint4 *)p, dst[i] = (uint4 *)p;
            }
            copy8(src, dst);
        }
    }
}

int main() {
    const int    N = 1250;
    const size_t image_bytes = 6266880; // 1080p 4:2:0 planes
    const int    blocks_per_image = (int)(image_bytes / 128);
    // list: runs of 12 listed blocks, 18 skipped (about 40 % listed, like the bench's G class), S a multiple of 32
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
    printf("device SMs %d, %d images x %d listed blocks = %.2f GB read + written per pass\n", sms, N, S, bytes / 1e9);
    auto run = [&](const char *name, int order, int ctas_per_sm, int threads, int window) {
        float best = 1e9f;
        for(int rep = 0; rep < 6; rep++) {
            cudaEventRecord(e0);
            if(order == 0) k_tile<<<sms * ctas_per_sm, threads>>>(base, image_bytes, dl, S, N);
            else k_op<<<sms * ctas_per_sm, threads>>>(base, image_bytes, dl, S, N, window);
            cudaEventRecord(e1);
            CK(cudaEventSynchronize(e1));
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            if(rep > 0 && ms < best) best = ms;
        }
        printf("%-40s ctas/SM %d threads %4d window %6d : %.3f ms  %.0f GB/s\n", name, ctas_per_sm, threads, window, best, bytes / best / 1e6);
    };
    for(int threads : {512}) {
        for(int cps : {2}) {
            run("tile order (fp32 kernel)", 0, cps, threads, 0);
            run("operator order", 1, cps, threads, 0);
        }
    }
    {
        // the same two orders on the same 7.8 GB seen as 59 765 "images" of 128 KB: a warp's 32 rows then span 4 MB (2-3 pages of
        // 2 MB) instead of 200 MB (32 pages) -- separates address-translation cost from DRAM page locality
        const size_t ib2 = 131072;
        const int    N2 = (int)(image_bytes * N / ib2), S2 = 384;
        const double bytes2 = (double)S2 * N2 * 256.0;
        for(int order = 0; order < 2; order++) {
            float best = 1e9f;
            for(int rep = 0; rep < 5; rep++) {
                cudaEventRecord(e0);
                if(order == 0) k_tile<<<sms * 2, 512>>>(base, ib2, dl, S2, N2);
                else k_op<<<sms * 2, 512>>>(base, ib2, dl, S2, N2, 0);
                cudaEventRecord(e1);
                CK(cudaEventSynchronize(e1));
                float ms;
                cudaEventElapsedTime(&ms, e0, e1);
                if(rep > 0 && ms < best) best = ms;
            }
            printf("128 KB images, %s order: %.3f ms  %.0f GB/s\n", order ? "operator" : "tile", best, bytes2 / best / 1e6);
        }
    }
    auto run3 = [&](int threads, int R) {
        float best = 1e9f;
        for(int rep = 0; rep < 6; rep++) {
            cudaEventRecord(e0);
            k_runs<<<sms, threads>>>(base, image_bytes, dl, S, N, R);
            cudaEventRecord(e1);
            CK(cudaEventSynchronize(e1));
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            if(rep > 0 && ms < best) best = ms;
        }
        printf("operator order, runs of %2d adjacent entries, %d threads : %.3f ms  %.0f GB/s\n", R, threads, best, bytes / best / 1e6);
    };
    for(int R : {1, 2, 4, 8, 16, 32}) {
        float best = 1e9f;
        for(int rep = 0; rep < 6; rep++) {
            cudaEventRecord(e0);
            k_adj<<<sms * 2, 512>>>(base, image_bytes, dl, S, N, R);
            cudaEventRecord(e1);
            CK(cudaEventSynchronize(e1));
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            if(rep > 0 && ms < best) best = ms;
        }
        printf("operator order, %2d adjacent entries x %2d images per warp : %.3f ms  %.0f GB/s\n", R, 32 / R, best, bytes / best / 1e6);
    }
    return 0;
}
