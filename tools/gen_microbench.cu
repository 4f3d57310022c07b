// Where do the generator's clocks go?  Register-only replicas of its per-value work on one CTA per SM, 8 warps (2 per scheduler), 16 values per
// thread and iteration -- the shape of the sweep kernel's generator -- timed with clock64:
//   v0: the 7 FP64 instructions of the table-driven exp, table value taken as a constant
//   v1: + the integer work (flush-to-zero test, index / exponent arithmetic)
//   v2: + the table lookup in shared memory (random indices)
//   v3: + the two dot-product DMMAs per 8 x 8 tile and the DADD that seeds them
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/gen_microbench tools/gen_microbench.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int V>
__global__ void __launch_bounds__(256, 1) k(double* out, long long* clk, int iters, double seed) {
    __shared__ double tab[2048];
    for (int i = threadIdx.x; i < 2048; i += 256) tab[i] = exp2((double)i / 2048.0);
    __syncthreads();
    const double MAGIC = 6755399441055744.0, C1 = 3.384507717577858e-04, C2 = 5.72744624517204e-08, C3 = 6.461528672932365e-12;
    double t[16], acc = 0.0;
    for (int c = 0; c < 16; ++c) t[c] = -seed * (threadIdx.x * 16 + c + 1);
    double xa = seed * 0.01 * threadIdx.x, zb = seed * 0.02 * threadIdx.x;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        double q[16], r[16], res[16]; int n[16];
        if (V >= 3) {
#pragma unroll
            for (int c = 0; c < 16; c += 2) { t[c] += zb; t[c + 1] += zb; }
#pragma unroll
            for (int c = 0; c < 16; c += 2) dmma(t[c], t[c + 1], xa, zb);
#pragma unroll
            for (int c = 0; c < 16; c += 2) dmma(t[c], t[c + 1], zb, xa);
        }
#pragma unroll
        for (int c = 0; c < 16; ++c) q[c] = t[c] + MAGIC;
#pragma unroll
        for (int c = 0; c < 16; ++c) {
            if (V >= 1) { const bool tiny = (unsigned)__double2hiint(t[c]) > 0xC13E8480u; n[c] = tiny ? (int)0x80000000 : __double2loint(q[c]); }
            else n[c] = 0;
            q[c] = q[c] - MAGIC;
        }
#pragma unroll
        for (int c = 0; c < 16; ++c) r[c] = t[c] - q[c];
#pragma unroll
        for (int c = 0; c < 16; ++c) q[c] = fma(r[c], C3, C2);
#pragma unroll
        for (int c = 0; c < 16; ++c) q[c] = fma(q[c], r[c], C1);
#pragma unroll
        for (int c = 0; c < 16; ++c) q[c] = q[c] * r[c];
#pragma unroll
        for (int c = 0; c < 16; ++c) {
            const double T = V >= 2 ? tab[n[c] & 2047] : 1.25;
            res[c] = fma(T, q[c], T);
            if (V >= 1) {
                const int hi = __double2hiint(res[c]) + ((n[c] >> 11) << 20);
                res[c] = __hiloint2double(hi, __double2loint(res[c]));
                if (n[c] == (int)0x80000000) res[c] = 0.0;
            }
        }
#pragma unroll
        for (int c = 0; c < 16; ++c) { acc += res[c]; t[c] = t[c] * 0.999 - 1.0e-3; }      // (2 more FP64 instructions per value: keeps the chain alive)
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
    out[blockIdx.x * 256 + threadIdx.x] = acc;
}
int main() {
    double* out; long long* clk; cudaMalloc(&out, 148 * 256 * 8); cudaMalloc(&clk, 148 * 8);
    const int iters = 2000;
    long long h[148];
    auto run = [&](auto kern, const char* name, int fp64_per_value) {
        kern<<<148, 256>>>(out, clk, iters, 1.0); cudaDeviceSynchronize();
        kern<<<148, 256>>>(out, clk, iters, 1.0); cudaDeviceSynchronize();
        cudaMemcpy(h, clk, sizeof h, cudaMemcpyDeviceToHost);
        double c = (double)h[0] / iters;                                   // clocks per iteration = 16 values x 256 threads = 4096 values per CTA
        printf("%-58s %8.0f clocks per 4096 values  (FP64 pipe time of the counted instructions: %d)\n", name, c, fp64_per_value * 4096 / 64);
    };
    run(k<0>, "v0 exp FP64 only (7 + 2 chain-keeping FP64 per value)", 9);
    run(k<1>, "v1 + integer work", 9);
    run(k<2>, "v2 + table lookup (smem, random index)", 9);
    run(k<3>, "v3 + DADD seed and two DMMAs per 8x8 tile", 10 + 8);
    printf("(sweep kernel's generator: ~2100 clocks per 4096 values in steady state; pipe time 1088)\n");
    return 0;
}
