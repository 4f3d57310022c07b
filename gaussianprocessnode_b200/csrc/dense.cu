// Small deterministic reductions of the M x M path, and the GEMM entry point the other translation units call.
//
// Replaces what the reference reaches per data point through LinearAlgebra:
//   meta.KuuL \ k, jdotavx(alpha, alpha)   GPnode/UniSGPnode.jl:208-209   -> sum_n = <K_uu^-1, Psi2>
//   mul!(k, meta.Uv, k), jdotavx(k, k)     GPnode/UniSGPnode.jl:212-213   -> sum_n = <R_v, Psi2>,  R_v = Sigma_v + mu_v mu_v' = Uv' Uv
// (the Cholesky / inverse kernels live in dense_coop.cu).  All matrices are column-major with leading dimension M.
#include "sgp_internal.cuh"
#include <cmath>
#include <algorithm>

namespace {

// out[0] = sum_i a[i*sa] * (b ? b[i*sb] : 1): block partials in a fixed order -> deterministic
__global__ void __launch_bounds__(256) dot_partial_kernel(const double* __restrict__ a, size_t sa, const double* __restrict__ b, size_t sb, size_t n,
                                                          double* __restrict__ partial) {
    __shared__ double s[256];
    double v0 = 0.0, v1 = 0.0;
    const size_t per = (n + gridDim.x - 1) / gridDim.x, lo = blockIdx.x * per, hi = lo + per < n ? lo + per : n;
    size_t i = lo + threadIdx.x;
    for (; i + 256 < hi; i += 512) {
        v0 = fma(a[i * sa], b ? b[i * sb] : 1.0, v0);
        v1 = fma(a[(i + 256) * sa], b ? b[(i + 256) * sb] : 1.0, v1);
    }
    if (i < hi) v0 = fma(a[i * sa], b ? b[i * sb] : 1.0, v0);
    s[threadIdx.x] = v0 + v1;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) { if (threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o]; __syncthreads(); }
    if (threadIdx.x == 0) partial[blockIdx.x] = s[0];
}
__global__ void dot_finish_kernel(const double* __restrict__ partial, int nb, double* __restrict__ out) {
    if (threadIdx.x == 0) { double v = 0.0; for (int i = 0; i < nb; ++i) v += partial[i]; out[0] = v; }
}

// One pass over the M x M matrices for every scalar the :w rule / energy / theta objective need.  Block b covers a contiguous range of
// (column, row part) tasks; partial[b][4]; the LAST block to finish (ticket counter) adds the partials in a fixed order: deterministic.
constexpr int WT_BLOCKS = 256;
__global__ void __launch_bounds__(256) wterms_kernel(const double* __restrict__ Kinv, const double* __restrict__ Psi2, const double* __restrict__ R,
                                                     const double* __restrict__ mu_outer, const double* __restrict__ mu, const double* __restrict__ psi1,
                                                     int M, int nparts, double* __restrict__ partial, unsigned* __restrict__ ticket, double* __restrict__ out) {
    __shared__ double s[4][8];
    __shared__ bool last;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // tasks = (column, row part): nparts row parts per column so that a small M still fills the machine with warps; a block takes a contiguous
    // range of tasks, a warp every eighth of them
    const int ntask = M * nparts;
    const int t_lo = (int)((long long)ntask * blockIdx.x / gridDim.x), t_hi = (int)((long long)ntask * (blockIdx.x + 1) / gridDim.x);
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    for (int t = t_lo + warp; t < t_hi; t += 8) {
        const int c = t / nparts, part = t - c * nparts;
        const int r_lo = (int)((long long)M * part / nparts), r_hi = (int)((long long)M * (part + 1) / nparts);
        const size_t off = (size_t)c * M;
        const double mc = mu_outer ? mu_outer[c] : 0.0;
#pragma unroll 4                                                 // (the loads of four iterations in flight)
        for (int r = r_lo + lane; r < r_hi; r += 32) {
            const double p2 = Psi2[off + r];
            if (Kinv) a0 = fma(Kinv[off + r], p2, a0);
            double rv = R ? R[off + r] : 0.0;
            if (mu_outer) rv = fma(mu_outer[r], mc, rv);
            a1 = fma(rv, p2, a1);
        }
        if (lane == 0 && part == 0) {
            if (mu && psi1) a2 = fma(mu[c], psi1[c], a2);
            if (Kinv) a3 += Kinv[off + c];
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a0 += __shfl_xor_sync(0xffffffffu, a0, o); a1 += __shfl_xor_sync(0xffffffffu, a1, o);
        a2 += __shfl_xor_sync(0xffffffffu, a2, o); a3 += __shfl_xor_sync(0xffffffffu, a3, o);
    }
    if (lane == 0) { s[0][warp] = a0; s[1][warp] = a1; s[2][warp] = a2; s[3][warp] = a3; }
    __syncthreads();
    if (tid < 4) {
        double v = 0.0;
        for (int q = 0; q < 8; ++q) v += s[tid][q];
        partial[4 * blockIdx.x + tid] = v;
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    __syncthreads();
    if (last) {
        __threadfence();
        if (warp < 4) {                                        // a warp per scalar: strided partial sums, then a fixed-shape tree
            double v = 0.0;
            for (unsigned b = lane; b < gridDim.x; b += 32) v += __ldcg(partial + 4 * b + warp);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0) out[warp] = v;
        }
        if (tid == 0) *ticket = 0u;                            // ready for the next call on this stream
    }
}

}  // namespace

int sgp_dot(sgp_ctx* ctx, const double* a, size_t sa, const double* b, size_t sb, size_t n, double* out) {
    const int nb = (int)std::min<size_t>(128, (n + 4095) / 4096);
    double* partial = ctx->dense_dev + SGP_MAX_D;            // [SGP_MAX_D, 64) of the scratch header holds up to 32 block partials
    const int nblk = nb > 32 ? 32 : (nb < 1 ? 1 : nb);
    dot_partial_kernel<<<nblk, 256, 0, ctx->stream>>>(a, sa, b, sb, n, partial);
    dot_finish_kernel<<<1, 32, 0, ctx->stream>>>(partial, nblk, out);
    SGP_CUDA(ctx, cudaGetLastError());
    return SGP_OK;
}

int sgp_wterms_reduce(sgp_ctx* ctx, const double* Kinv, const double* Psi2, const double* R, const double* mu_outer, const double* mu,
                      const double* psi1, int M, double* out) {
    if (!ctx->wt_dev) {
        SGP_CUDA(ctx, cudaMalloc((void**)&ctx->wt_dev, (4 * WT_BLOCKS + 8) * sizeof(double)));
        SGP_CUDA(ctx, cudaMemsetAsync(ctx->wt_dev, 0, (4 * WT_BLOCKS + 8) * sizeof(double), ctx->stream));
    }
    const int nparts = std::max(1, std::min(8, 2048 / std::max(1, M)));
    const int nblk = std::max(1, std::min(WT_BLOCKS, M * nparts / 8));
    wterms_kernel<<<nblk, 256, 0, ctx->stream>>>(Kinv, Psi2, R, mu_outer, mu, psi1, M, nparts, ctx->wt_dev, reinterpret_cast<unsigned*>(ctx->wt_dev + 4 * WT_BLOCKS), out);
    SGP_CUDA(ctx, cudaGetLastError());
    return SGP_OK;
}

int sgp_gemm2(sgp_ctx* ctx, int opA, int opB, int m, int n, int k, double alpha, const double* A, int lda, const double* B, int ldb, double beta,
              double* C, int ldc, int lower_only);

int sgp_gemm(sgp_ctx* ctx, int opA, int opB, int m, int n, int k, double alpha, const double* A, int lda, const double* B, int ldb,
             double beta, double* C, int ldc, int lower_only) {
    return sgp_gemm2(ctx, opA, opB, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, lower_only);
}
