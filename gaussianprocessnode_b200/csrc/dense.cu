// M x M factorisations of the path, hand-written for sm_100a: blocked right-looking Cholesky, blocked triangular solves
// and a DMMA (mma.sync m8n8k4 f64) tile GEMM for the trailing updates / posterior products.
//
// Replaces the LAPACK/BLAS calls the reference reaches through LinearAlgebra and FastCholesky (unvendored):
//   fastcholesky!(Kuu).L           experiments/regression_kin40k.ipynb:183-184, classification_banana.ipynb:163-164
//   cholinv(Kuu), inv(Kuu)         experiments/Pendulum_Wishart_2d.ipynb:2542-2543, GPLVM.ipynb:248-249
//   mean_cov(marginal_v) (cholinv of the precision), mul!(Sigma_v, mu_v, mu_v', 1, 1), fastcholesky!(Sigma_v).U
//                                  GPnode/UniSGPnode.jl:62-73
//   meta.KuuL \ k, meta.Uv * k     GPnode/UniSGPnode.jl:208-213 (per point; here once per sweep on Psi2)
// All matrices are column-major with leading dimension M.  Sizes on this path are M = 20 ... 1024 (D*M <= ~2000 for
// MultiSGP), i.e. latency-bound: the kernels are organised as few launches per 64-wide panel, not as a FLOP race.
#include "sgp_internal.cuh"
#include <cmath>
#include <algorithm>
#include <cstdlib>

namespace {

constexpr int PB = 64;   // panel width

// ---- generic tile GEMM: C[m x n] = beta*C + alpha * op(A) op(B), column-major, DMMA ------------------------------
// opA: 0 -> A is m x k (lda), 1 -> A is k x m (use A').  opB: 0 -> B is k x n, 1 -> B is n x k (use B').
// lower_only: skip tiles strictly above the diagonal (C symmetric / triangular updates).
struct GemmArgs {
    const double* A; const double* B; double* C;
    int m, n, k, lda, ldb, ldc, opA, opB, lower_only;
    double alpha, beta;
};

__global__ void __launch_bounds__(128) gemm_kernel(const GemmArgs g) {
    constexpr int T = 64, KT = 16, LDS = T + 4;
    __shared__ double As[KT * LDS];   // As[kk][mm]
    __shared__ double Bs[KT * LDS];   // Bs[kk][nn]
    const int bm = blockIdx.x * T, bn = blockIdx.y * T;
    if (g.lower_only && bn > bm + T - 1) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wr = warp >> 1, wc = warp & 1;          // 2 x 2 warps, warp tile 32 x 32
    double acc[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    for (int k0 = 0; k0 < g.k; k0 += KT) {
        // stage A tile: element (mm, kk) for mm < T, kk < KT
        for (int e = tid; e < T * KT; e += 128) {
            int mm, kk;
            if (g.opA == 0) { mm = e % T; kk = e / T; } else { kk = e % KT; mm = e / KT; }
            int gm = bm + mm, gk = k0 + kk;
            double v = 0.0;
            if (gm < g.m && gk < g.k) v = g.opA == 0 ? g.A[(size_t)gm + (size_t)gk * g.lda] : g.A[(size_t)gk + (size_t)gm * g.lda];
            As[kk * LDS + mm] = v;
        }
        for (int e = tid; e < T * KT; e += 128) {
            int nn, kk;
            if (g.opB == 0) { kk = e % KT; nn = e / KT; } else { nn = e % T; kk = e / T; }
            int gn = bn + nn, gk = k0 + kk;
            double v = 0.0;
            if (gn < g.n && gk < g.k) v = g.opB == 0 ? g.B[(size_t)gk + (size_t)gn * g.ldb] : g.B[(size_t)gn + (size_t)gk * g.ldb];
            Bs[kk * LDS + nn] = v;
        }
        __syncthreads();
#pragma unroll
        for (int ks = 0; ks < KT / 4; ++ks) {
            const double* ar = As + (ks * 4 + (lane & 3)) * LDS + wr * 32 + (lane >> 2);
            const double* br = Bs + (ks * 4 + (lane & 3)) * LDS + wc * 32 + (lane >> 2);
            double a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { a[i] = ar[8 * i]; b[i] = br[8 * i]; }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                int gm = bm + wr * 32 + 8 * i + (lane >> 2), gn = bn + wc * 32 + 8 * j + 2 * (lane & 3) + h;
                if (gm < g.m && gn < g.n) {
                    double* c = g.C + (size_t)gm + (size_t)gn * g.ldc;
                    double old = g.beta == 0.0 ? 0.0 : g.beta * *c;
                    *c = fma(g.alpha, acc[i][j][h], old);
                }
            }
}

int gemm(sgp_ctx* ctx, int opA, int opB, int m, int n, int k, double alpha, const double* A, int lda, const double* B, int ldb,
         double beta, double* C, int ldc, int lower_only = 0) {
    if (m <= 0 || n <= 0) return SGP_OK;
    GemmArgs g{A, B, C, m, n, k, lda, ldb, ldc, opA, opB, lower_only, alpha, beta};
    dim3 grid((m + 63) / 64, (n + 63) / 64);
    gemm_kernel<<<grid, 128, 0, ctx->stream>>>(g);
    SGP_CUDA(ctx, cudaGetLastError());
    return SGP_OK;
}

// (the Cholesky itself lives in dense_coop.cu: one cooperative kernel per factorisation)

__global__ void __launch_bounds__(64) trsm_diag_kernel(const double* Lkk, int ldl, double* B, int ldb, int nb, int nrhs, int trans) {
    __shared__ double L[PB][PB + 1];
    for (int e = threadIdx.x; e < nb * nb; e += 64) { int r = e % nb, c = e / nb; L[r][c] = Lkk[(size_t)r + (size_t)c * ldl]; }
    __syncthreads();
    int c = blockIdx.x * 64 + threadIdx.x;
    if (c >= nrhs) return;
    double* b = B + (size_t)c * ldb;
    double x[PB];
    if (!trans) {
#pragma unroll 1
        for (int j = 0; j < nb; ++j) {
            double v = b[j];
            for (int l = 0; l < j; ++l) v = fma(-L[j][l], x[l], v);
            x[j] = v / L[j][j];
        }
    } else {
#pragma unroll 1
        for (int j = nb - 1; j >= 0; --j) {
            double v = b[j];
            for (int l = j + 1; l < nb; ++l) v = fma(-L[l][j], x[l], v);
            x[j] = v / L[j][j];
        }
    }
    for (int j = 0; j < nb; ++j) b[j] = x[j];
}

__global__ void kuu_kernel(const double* __restrict__ Z, double* __restrict__ K, int M, int D, int kind, double variance, const double* ell_inv,
                           double jitter) {
    size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (e >= (size_t)M * M) return;
    int i = (int)(e % M), j = (int)(e / M);
    double r2 = 0.0;
    for (int d = 0; d < D; ++d) { double t = (Z[(size_t)i * D + d] - Z[(size_t)j * D + d]) * ell_inv[d]; r2 = fma(t, t, r2); }
    double v;
    if (kind == SGP_KERNEL_SE) v = variance * exp(-0.5 * r2);
    else if (kind == SGP_KERNEL_MATERN32) { double s = sqrt(3.0 * r2); v = variance * (1.0 + s) * exp(-s); }
    else { double s = sqrt(5.0 * r2); v = variance * (1.0 + s + s * s / 3.0) * exp(-s); }
    if (i == j) v += jitter;
    K[e] = v;
}



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

int sgp_gemm2(sgp_ctx* ctx, int opA, int opB, int m, int n, int k, double alpha, const double* A, int lda, const double* B, int ldb, double beta,
              double* C, int ldc, int lower_only);

int sgp_gemm(sgp_ctx* ctx, int opA, int opB, int m, int n, int k, double alpha, const double* A, int lda, const double* B, int ldb,
             double beta, double* C, int ldc, int lower_only) {
    if (std::getenv("SGP_GEMM_V1")) return gemm(ctx, opA, opB, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, lower_only);
    return sgp_gemm2(ctx, opA, opB, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, lower_only);
}

// B (M x nrhs, ld M) <- L^{-1} B (trans = false) or L^{-T} B (trans = true); L lower, column-major, ld M.
int sgp_trsm_lower(sgp_ctx* ctx, const double* L, double* B, int M, int nrhs, bool trans) {
    int nblk = (M + PB - 1) / PB;
    if (!trans) {
        for (int kb = 0; kb < nblk; ++kb) {
            int k = kb * PB, nb = M - k < PB ? M - k : PB;
            trsm_diag_kernel<<<(nrhs + 63) / 64, 64, 0, ctx->stream>>>(L + (size_t)k + (size_t)k * M, M, B + k, M, nb, nrhs, 0);
            int rows = M - k - nb;
            if (rows > 0) {   // B2 -= L21 X1
                int rc = gemm(ctx, 0, 0, rows, nrhs, nb, -1.0, L + (size_t)(k + nb) + (size_t)k * M, M, B + k, M, 1.0, B + k + nb, M);
                if (rc) return rc;
            }
        }
    } else {
        for (int kb = nblk - 1; kb >= 0; --kb) {
            int k = kb * PB, nb = M - k < PB ? M - k : PB;
            trsm_diag_kernel<<<(nrhs + 63) / 64, 64, 0, ctx->stream>>>(L + (size_t)k + (size_t)k * M, M, B + k, M, nb, nrhs, 1);
            if (k > 0) {      // B1 -= L21' X2   (L21 = L[k:k+nb, 0:k])
                int rc = gemm(ctx, 1, 0, k, nrhs, nb, -1.0, L + (size_t)k, M, B + k, M, 1.0, B, M);
                if (rc) return rc;
            }
        }
    }
    SGP_CUDA(ctx, cudaGetLastError());
    return SGP_OK;
}

int sgp_kuu_build(sgp_ctx* ctx, double* K, double jitter) {
    const int M = ctx->M, D = ctx->D;
    double inv[SGP_MAX_D];
    for (int d = 0; d < D; ++d) inv[d] = 1.0 / ctx->ell[d];
    double* ell_dev = ctx->dense_dev;   // first SGP_MAX_D doubles of the scratch are reserved for this
    SGP_CUDA(ctx, cudaMemcpyAsync(ell_dev, inv, sizeof(double) * D, cudaMemcpyHostToDevice, ctx->stream));
    kuu_kernel<<<(unsigned)(((size_t)M * M + 255) / 256), 256, 0, ctx->stream>>>(ctx->Z_dev, K, M, D, ctx->kind, ctx->variance, ell_dev, jitter);
    SGP_CUDA(ctx, cudaGetLastError());
    SGP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));   // inv[] is a stack buffer
    return SGP_OK;
}
