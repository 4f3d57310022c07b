// Dependent-issue latencies of the instructions on the critical path of the M x M kernels (one warp, chains of N dependent operations).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lat_microbench lat_microbench.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int MODE>
__global__ void lat(double* out, long long* clk, double seed, int n) {
    __shared__ double sh[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sh[i] = seed * i;
    __syncthreads();
    double x = seed + threadIdx.x * 1e-3, y = 0.0, c0 = 0.0, c1 = 0.0;
    int idx = threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) {
        if (MODE == 0) x = fma(x, 1.0000001, 1e-9);                                   // DFMA chain
        if (MODE == 1) x = __shfl_sync(0xffffffffu, x, (threadIdx.x + 1) & 31);       // shuffle of a double (2 x SHFL)
        if (MODE == 2) x = rsqrt(x) + 1.5;                                            // rsqrt + dadd
        if (MODE == 3) dmma(c0, c1, x, x);                                            // DMMA accumulator chain
        if (MODE == 4) { idx = (int)sh[idx & 1023] & 1023; }                          // LDS.64 + F2I chain
        if (MODE == 5) x = sqrt(x) + 1.5;
        if (MODE == 6) x = 1.0 / x + 1.5;
        if (MODE == 7) { x = fma(x, 1.0000001, 1e-9); y = fma(y, 1.0000001, x); }     // 2 DFMA, second depends on the first
        if (MODE == 8) x = sh[(threadIdx.x + i) & 1023] + x;                          // LDS (independent address) + DADD chain
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) clk[MODE] = t1 - t0;
    out[threadIdx.x + 32 * MODE] = x + y + c0 + c1 + idx;
}
int main() {
    double* out; long long* clk; cudaMalloc(&out, 4096 * 8); cudaMalloc(&clk, 16 * 8);
    const int n = 4096;
    const char* names[] = {"DFMA", "SHFL.f64", "rsqrt+dadd", "DMMA acc chain", "LDS+F2I", "sqrt+dadd", "rcp+dadd", "2xDFMA", "LDS+DADD"};
    for (int rep = 0; rep < 2; ++rep) {
        lat<0><<<1, 32>>>(out, clk, 1.0, n); lat<1><<<1, 32>>>(out, clk, 1.0, n); lat<2><<<1, 32>>>(out, clk, 1.0, n); lat<3><<<1, 32>>>(out, clk, 1.0, n);
        lat<4><<<1, 32>>>(out, clk, 1.0, n); lat<5><<<1, 32>>>(out, clk, 1.0, n); lat<6><<<1, 32>>>(out, clk, 1.0, n); lat<7><<<1, 32>>>(out, clk, 1.0, n);
        lat<8><<<1, 32>>>(out, clk, 1.0, n);
        cudaDeviceSynchronize();
    }
    long long h[16]; cudaMemcpy(h, clk, sizeof h, cudaMemcpyDeviceToHost);
    for (int m = 0; m < 9; ++m) printf("%-16s %.1f clocks per iteration\n", names[m], (double)h[m] / n);
    return 0;
}
