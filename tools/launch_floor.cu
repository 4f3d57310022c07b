// What a launch costs before the first and after the last instruction of a persistent kernel: [256 MB memset = L2 flush] [event] kernel [event] on
// one stream, as bench.py times a sweep, for an (almost) empty kernel of the sweep's launch shape -- 148 CTAs x 256 threads, 222 080 B of dynamic
// shared memory, cooperative, 255 registers' worth of occupancy (one CTA per SM) -- and for lighter shapes.  The difference between the sweep's
// event-timed duration (121 us) and its in-kernel clocks (108 us) is compared with these floors in DESIGN.md section 4.1.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/launch_floor tools/launch_floor.cu
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>
#include <algorithm>
namespace cg = cooperative_groups;

__global__ void __launch_bounds__(256, 1) k_coop(long long* out, int syncs) {
    extern __shared__ double sm[];
    const long long t0 = clock64();
    sm[threadIdx.x] = (double)t0;
    cg::grid_group g = cg::this_grid();
    for (int i = 0; i < syncs; ++i) g.sync();
    if (threadIdx.x == 0) out[blockIdx.x] = clock64() - t0 + (long long)sm[1] * 0;
}
__global__ void __launch_bounds__(256, 1) k_plain(long long* out) {
    extern __shared__ double sm[];
    const long long t0 = clock64();
    sm[threadIdx.x] = (double)t0;
    __syncthreads();
    if (threadIdx.x == 0) out[blockIdx.x] = clock64() - t0 + (long long)sm[1] * 0;
}

int main() {
    cudaStream_t st; cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
    void* flush; cudaMalloc(&flush, 256u << 20);
    long long* out; cudaMalloc(&out, 4096);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    struct Case { const char* name; int coop, smem, syncs, flushed; };
    const Case cases[] = {{"cooperative, 222080 B smem, 1 grid.sync, after L2 flush", 1, 222080, 1, 1}, {"cooperative, 222080 B smem, 0 grid.sync, after L2 flush", 1, 222080, 0, 1},
                          {"cooperative, 222080 B smem, 1 grid.sync, no flush", 1, 222080, 1, 0}, {"plain, 222080 B smem, after L2 flush", 0, 222080, 0, 1},
                          {"plain, 1024 B smem, after L2 flush", 0, 1024, 0, 1}, {"plain, 1024 B smem, no flush", 0, 1024, 0, 0}};
    cudaFuncSetAttribute(k_coop, cudaFuncAttributeMaxDynamicSharedMemorySize, 222080);
    cudaFuncSetAttribute(k_plain, cudaFuncAttributeMaxDynamicSharedMemorySize, 222080);
    for (const Case& c : cases) {
        std::vector<float> ms;
        for (int r = 0; r < 60; ++r) {
            if (c.flushed) cudaMemsetAsync(flush, r & 0xff, 256u << 20, st);
            cudaEventRecord(e0, st);
            int syncs = c.syncs;
            void* args[] = {&out, &syncs};
            if (c.coop) cudaLaunchCooperativeKernel((const void*)k_coop, dim3(148), dim3(256), args, c.smem, st);
            else k_plain<<<148, 256, c.smem, st>>>(out);
            cudaEventRecord(e1, st);
            cudaEventSynchronize(e1);
            float t = 0.f; cudaEventElapsedTime(&t, e0, e1);
            if (r >= 10) ms.push_back(t);
        }
        std::sort(ms.begin(), ms.end());
        double mean = 0; for (float t : ms) mean += t; mean /= ms.size();
        printf("%-62s event-to-event: median %.2f us, mean %.2f us, min %.2f us   (%s)\n", c.name, ms[ms.size() / 2] * 1e3, mean * 1e3, ms[0] * 1e3, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
