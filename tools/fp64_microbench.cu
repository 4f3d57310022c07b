// FP64 pipe microbenchmarks for B200 (sm_100a): DFMA vs DMMA issue rates, and whether
// they share a datapath. Output decides the fused K_uf->Psi2 kernel's tile plan.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_microbench fp64_microbench.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

__device__ __forceinline__ void mma884(double (&c)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}
__device__ __forceinline__ void mma1684(double (&c)[4], const double (&a)[2], double b) {
    asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a[0]), "d"(a[1]), "d"(b));
}
__device__ __forceinline__ void mma1688(double (&c)[4], const double (&a)[4], const double (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void mma16816(double (&c)[4], const double (&a)[8], const double (&b)[4]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                   "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

// mode 0: DFMA only (NACC chains). mode 1: m8n8k4. 2: m16n8k4. 3: m16n8k8. 4: m16n8k16.
// mode 5: m16n8k16 + NF DFMA per MMA in the same warp. mode 6: even warps MMA, odd warps DFMA.
template <int MODE, int NACC, int NF>
__global__ void __launch_bounds__(MODE <= 1 ? 1024 : 512) bench(double* out, int iters, double seed) {
    double c[NACC][4];
    double f[16];
    double a8[8], b4[4];
#pragma unroll
    for (int i = 0; i < NACC; i++) { c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.0; }
#pragma unroll
    for (int i = 0; i < 16; i++) f[i] = seed * (i + threadIdx.x);
#pragma unroll
    for (int i = 0; i < 8; i++) a8[i] = seed * (threadIdx.x + i) * 1e-3;
#pragma unroll
    for (int i = 0; i < 4; i++) b4[i] = seed * (threadIdx.x - i) * 1e-3;
    const double m = 1.0 + seed * 1e-9, ad = seed * 1e-7;
    const bool mma_warp = (MODE != 6) || ((threadIdx.x >> 5) & 1) == 0;
    for (int it = 0; it < iters; it++) {
        if (MODE == 0 || (MODE == 6 && !mma_warp)) {
#pragma unroll
            for (int r = 0; r < 4; r++)
#pragma unroll
                for (int i = 0; i < 16; i++) f[i] = fma(f[i], m, ad);
        }
        if (MODE == 1) {
#pragma unroll
            for (int i = 0; i < NACC; i++) { double cc[2] = {c[i][0], c[i][1]}; mma884(cc, a8[i & 7], b4[i & 3]); c[i][0] = cc[0]; c[i][1] = cc[1]; }
        }
        if (MODE == 2) {
#pragma unroll
            for (int i = 0; i < NACC; i++) { double aa[2] = {a8[i & 7], a8[(i + 1) & 7]}; mma1684(c[i], aa, b4[i & 3]); }
        }
        if (MODE == 3) {
#pragma unroll
            for (int i = 0; i < NACC; i++) { double aa[4] = {a8[i & 7], a8[(i + 1) & 7], a8[(i + 2) & 7], a8[(i + 3) & 7]}; double bb[2] = {b4[i & 3], b4[(i + 1) & 3]}; mma1688(c[i], aa, bb); }
        }
        if (MODE == 4 || MODE == 5 || (MODE == 6 && mma_warp)) {
#pragma unroll
            for (int i = 0; i < NACC; i++) {
                mma16816(c[i], a8, b4);
                if (MODE == 5) {
#pragma unroll
                    for (int k = 0; k < NF; k++) f[k & 15] = fma(f[k & 15], m, ad);
                }
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; i++) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
#pragma unroll
    for (int i = 0; i < 16; i++) s += f[i];
    if (s == 123.456) out[0] = s;
}

template <int MODE, int NACC, int NF>
void run(const char* name, int threads, int blocks_per_sm, double fma_per_thread_iter, double extra_fma_per_thread_iter, double* d_out, int nsm) {
    int iters = 20000;
    int grid = nsm * blocks_per_sm;
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    bench<MODE, NACC, NF><<<grid, threads>>>(d_out, 200, 1.0);
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int rep = 0; rep < 3; rep++) {
        CK(cudaEventRecord(e0));
        bench<MODE, NACC, NF><<<grid, threads>>>(d_out, iters, 1.0);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    double total_threads = (double)grid * threads;
    double mma_flops = 2.0 * fma_per_thread_iter * iters * total_threads;
    double extra = 2.0 * extra_fma_per_thread_iter * iters * total_threads;
    printf("%-44s thr=%4d blk/sm=%d  %8.3f ms  main %7.2f TF  extra %7.2f TF  sum %7.2f TF\n", name, threads, blocks_per_sm, best,
           mma_flops / best * 1e-9, extra / best * 1e-9, (mma_flops + extra) / best * 1e-9);
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int nsm = p.multiProcessorCount;
    printf("device %s, %d SMs, clock %d kHz\n", p.name, nsm, p.clockRate);
    double* d_out; CK(cudaMalloc(&d_out, 64));
    // per-thread FMA per iteration: DFMA mode = 64 ; m8n8k4 = 256/32 = 8 per mma ; m16n8k4 = 16 ; m16n8k8 = 32 ; m16n8k16 = 64
    for (int thr : {128, 256, 512, 1024}) {
        run<0, 1, 0>("DFMA x64/iter", thr, 1, 64, 0, d_out, nsm);
    }
    for (int thr : {128, 256, 512, 1024}) {
        run<1, 8, 0>("DMMA m8n8k4 x8 acc", thr, 1, 8 * 8, 0, d_out, nsm);
    }
    for (int thr : {128, 256, 512}) {
        run<2, 8, 0>("DMMA m16n8k4 x8 acc", thr, 1, 16 * 8, 0, d_out, nsm);
        run<3, 8, 0>("DMMA m16n8k8 x8 acc", thr, 1, 32 * 8, 0, d_out, nsm);
        run<4, 8, 0>("DMMA m16n8k16 x8 acc", thr, 1, 64 * 8, 0, d_out, nsm);
        run<4, 4, 0>("DMMA m16n8k16 x4 acc", thr, 1, 64 * 4, 0, d_out, nsm);
        run<4, 2, 0>("DMMA m16n8k16 x2 acc", thr, 1, 64 * 2, 0, d_out, nsm);
        run<4, 1, 0>("DMMA m16n8k16 x1 acc (latency)", thr, 1, 64 * 1, 0, d_out, nsm);
    }
    for (int thr : {256, 512}) {
        run<5, 8, 4>("m16n8k16 + 4 DFMA/mma  (+6%)", thr, 1, 64 * 8, 4 * 8, d_out, nsm);
        run<5, 8, 8>("m16n8k16 + 8 DFMA/mma  (+12%)", thr, 1, 64 * 8, 8 * 8, d_out, nsm);
        run<5, 8, 16>("m16n8k16 + 16 DFMA/mma (+25%)", thr, 1, 64 * 8, 16 * 8, d_out, nsm);
        run<5, 8, 32>("m16n8k16 + 32 DFMA/mma (+50%)", thr, 1, 64 * 8, 32 * 8, d_out, nsm);
        run<5, 8, 64>("m16n8k16 + 64 DFMA/mma (+100%)", thr, 1, 64 * 8, 64 * 8, d_out, nsm);
        // mode 6: half warps mma (64*8 per thread-iter on half the threads), half DFMA (64 per iter)
        run<6, 8, 0>("split warps: even MMA(k16) / odd DFMA", thr, 1, 64 * 8 * 0.5, 64 * 0.5, d_out, nsm);
    }
    return 0;
}
