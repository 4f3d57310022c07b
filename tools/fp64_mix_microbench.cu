// FP64 pipe microbenchmark 2 (B200, sm_100a): how DMMA.8x8x4 and DFMA share the pipe as a function of the interleave
// granularity, and how the scheduler arbitrates between a DMMA warp and a DFMA warp on the same sub-partition.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_mix_microbench fp64_mix_microbench.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

__device__ __forceinline__ void mma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dfma(double& f, double m, double a) { asm volatile("fma.rn.f64 %0, %0, %1, %2;\n" : "+d"(f) : "d"(m), "d"(a)); }

// MODE 0: every warp: groups of G DMMAs followed by G*R DFMAs (16 independent chains), 32 DMMAs per iteration
// MODE 1: warps 0..3 (one per sub-partition) DMMA only, warps 4..7 DFMA only; each counts its own iterations until the
//         other side has finished a fixed amount -> reports both rates
template <int G, int R>
__global__ void __launch_bounds__(256) mix(double* out, int iters, double seed) {
    double c[32][2], f[16];
#pragma unroll
    for (int i = 0; i < 32; ++i) c[i][0] = c[i][1] = 0.0;
#pragma unroll
    for (int i = 0; i < 16; ++i) f[i] = seed * (i + threadIdx.x);
    const double a = seed * threadIdx.x * 1e-3, b = seed * 1e-3, m = 1.0 + seed * 1e-9, ad = seed * 1e-7;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int g0 = 0; g0 < 32; g0 += G) {
#pragma unroll
            for (int g = 0; g < G; ++g) mma884(c[g0 + g][0], c[g0 + g][1], a, b);
#pragma unroll
            for (int k = 0; k < G * R; ++k) dfma(f[(g0 * R + k) & 15], m, ad);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 32; ++i) s += c[i][0] + c[i][1];
#pragma unroll
    for (int i = 0; i < 16; ++i) s += f[i];
    if (s == 123.456) out[0] = s;
}

__global__ void __launch_bounds__(256) split(double* out, long long* cnt, int iters, double seed, int nmma_warps_per_smsp) {
    double c[32][2], f[16];
#pragma unroll
    for (int i = 0; i < 32; ++i) c[i][0] = c[i][1] = 0.0;
#pragma unroll
    for (int i = 0; i < 16; ++i) f[i] = seed * (i + threadIdx.x);
    const double a = seed * threadIdx.x * 1e-3, b = seed * 1e-3, m = 1.0 + seed * 1e-9, ad = seed * 1e-7;
    const int warp = threadIdx.x >> 5;
    __shared__ volatile int done;
    if (threadIdx.x == 0) done = 0;
    __syncthreads();
    long long n = 0;
    if (warp < 4 * nmma_warps_per_smsp) {          // DMMA warps: fixed amount of work
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int g = 0; g < 32; ++g) mma884(c[g][0], c[g][1], a, b);
        }
        if (threadIdx.x == 0) done = 1;
    } else {                                       // DFMA warps: run until the DMMA warps are done, count iterations
        while (!done) {
#pragma unroll
            for (int k = 0; k < 64; ++k) dfma(f[k & 15], m, ad);
            ++n;
        }
        if ((threadIdx.x & 31) == 0) cnt[blockIdx.x * 8 + warp] = n;
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 32; ++i) s += c[i][0] + c[i][1];
#pragma unroll
    for (int i = 0; i < 16; ++i) s += f[i];
    if (s == 123.456) out[0] = s;
}

template <int G, int R>
void run(double* d_out, int nsm, int threads) {
    const int iters = 4000;
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    mix<G, R><<<nsm, threads>>>(d_out, 100, 1.0); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        CK(cudaEventRecord(e0)); mix<G, R><<<nsm, threads>>>(d_out, iters, 1.0); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    double warps = (double)nsm * threads / 32;
    double mma = 2.0 * 256 * 32 * iters * warps, fma = 2.0 * 32 * 32 * R * iters * warps;
    double clk_per_iter_smsp = best * 1e-3 * 1.965e9 / iters / (threads / 128.0);   // per warp-iteration of 32 DMMA + 32R DFMA
    printf("G=%2d DMMA then %2d DFMA (ratio %d DFMA/DMMA) thr=%d: %7.3f ms  mma %6.2f TF + fma %5.2f TF = %6.2f TF ; %.0f clk per (32 DMMA + %d DFMA), ideal %d\n",
           G, G * R, R, threads, best, mma / best * 1e-9, fma / best * 1e-9, (mma + fma) / best * 1e-9, clk_per_iter_smsp, 32 * R, 32 * 16 + 32 * R * 2);
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int nsm = p.multiProcessorCount;
    double* d_out; CK(cudaMalloc(&d_out, 64));
    long long* d_cnt; CK(cudaMalloc(&d_cnt, nsm * 8 * sizeof(long long)));
    for (int thr : {256, 512}) {
        run<32, 2>(d_out, nsm, thr); run<8, 2>(d_out, nsm, thr); run<4, 2>(d_out, nsm, thr); run<2, 2>(d_out, nsm, thr); run<1, 2>(d_out, nsm, thr);
        run<32, 1>(d_out, nsm, thr); run<1, 1>(d_out, nsm, thr); run<1, 4>(d_out, nsm, thr); run<32, 4>(d_out, nsm, thr);
    }
    for (int nm : {1}) {
        const int iters = 4000;
        cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        CK(cudaMemset(d_cnt, 0, nsm * 8 * sizeof(long long)));
        split<<<nsm, 256>>>(d_out, d_cnt, 100, 1.0, nm); CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(e0)); split<<<nsm, 256>>>(d_out, d_cnt, iters, 1.0, nm); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        long long h[8]; CK(cudaMemcpy(h, d_cnt, sizeof(h), cudaMemcpyDeviceToHost));
        double mma = 2.0 * 256 * 32 * iters * nsm * 4 * nm;
        double fma = 2.0 * 32 * 64 * (double)(h[4] + h[5] + h[6] + h[7]) * nsm;
        printf("split: %d DMMA warp + 1 DFMA warp per sub-partition: %.3f ms  mma %.2f TF, fma %.2f TF (DFMA warp iterations %lld)\n", nm, ms,
               mma / ms * 1e-9, fma / ms * 1e-9, h[4]);
    }
    return 0;
}
