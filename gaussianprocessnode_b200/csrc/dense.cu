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

// ---- Cholesky -----------------------------------------------------------------------------------------------------
// Diagonal block (nb <= 64) factorised by one CTA in shared memory; info = first non-positive pivot (1-based) or 0.
__global__ void __launch_bounds__(256) potrf_diag_kernel(double* A, int lda, int nb, int offset, int* info) {
    __shared__ double S[PB][PB + 1];
    const int tid = threadIdx.x;
    for (int e = tid; e < nb * nb; e += 256) { int r = e % nb, c = e / nb; S[r][c] = A[(size_t)r + (size_t)c * lda]; }
    __syncthreads();
    for (int j = 0; j < nb; ++j) {
        double d = S[j][j];
        if (!(d > 0.0)) {   // also catches NaN
            if (tid == 0 && *info == 0) *info = offset + j + 1;
            return;
        }
        double sd = sqrt(d);
        __syncthreads();
        if (tid == 0) S[j][j] = sd;
        for (int r = j + 1 + tid; r < nb; r += 256) S[r][j] /= sd;
        __syncthreads();
        // trailing update of the remaining columns: S[r][c] -= S[r][j] * S[c][j]  (c > j, r >= c)
        int rem = nb - j - 1;
        for (int e = tid; e < rem * rem; e += 256) {
            int r = j + 1 + e % rem, c = j + 1 + e / rem;
            if (r >= c) S[r][c] = fma(-S[r][j], S[c][j], S[r][c]);
        }
        __syncthreads();
    }
    for (int e = tid; e < nb * nb; e += 256) {
        int r = e % nb, c = e / nb;
        A[(size_t)r + (size_t)c * lda] = (r >= c) ? S[r][c] : 0.0;    // explicit zeros above the diagonal
    }
}

// Panel below the diagonal block: X L_kk' = A  (row-wise forward substitution), one thread per row.
__global__ void __launch_bounds__(64) potrf_panel_kernel(const double* Lkk, double* P, int lda, int nb, int rows, const int* info) {
    if (*info != 0) return;
    __shared__ double L[PB][PB + 1];
    for (int e = threadIdx.x; e < nb * nb; e += 64) { int r = e % nb, c = e / nb; L[r][c] = Lkk[(size_t)r + (size_t)c * lda]; }
    __syncthreads();
    int r = blockIdx.x * 64 + threadIdx.x;
    if (r >= rows) return;
    double x[PB];
#pragma unroll 1
    for (int j = 0; j < nb; ++j) {
        double v = P[(size_t)r + (size_t)j * lda];
        for (int l = 0; l < j; ++l) v = fma(-x[l], L[j][l], v);
        x[j] = v / L[j][j];
    }
    for (int j = 0; j < nb; ++j) P[(size_t)r + (size_t)j * lda] = x[j];
}

__global__ void zero_upper_kernel(double* A, int M) {
    size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (e >= (size_t)M * M) return;
    int r = (int)(e % M), c = (int)(e / M);
    if (r < c) A[e] = 0.0;
}

// ---- triangular solves with many right-hand sides ---------------------------------------------------------------
// Diagonal block solve: one thread per right-hand-side column.  trans = 0: L x = b (forward); 1: L' x = b (backward).
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

}  // namespace

int sgp_gemm(sgp_ctx* ctx, int opA, int opB, int m, int n, int k, double alpha, const double* A, int lda, const double* B, int ldb,
             double beta, double* C, int ldc, int lower_only) {
    return gemm(ctx, opA, opB, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, lower_only);
}

// In-place lower Cholesky of the column-major M x M matrix A (upper triangle is zeroed).
int sgp_potrf_lower(sgp_ctx* ctx, double* A, int M) {
    SGP_CUDA(ctx, cudaMemsetAsync(ctx->info_dev, 0, sizeof(int), ctx->stream));
    for (int k = 0; k < M; k += PB) {
        int nb = M - k < PB ? M - k : PB;
        double* Akk = A + (size_t)k + (size_t)k * M;
        potrf_diag_kernel<<<1, 256, 0, ctx->stream>>>(Akk, M, nb, k, ctx->info_dev);
        int rows = M - k - nb;
        if (rows > 0) {
            double* P = A + (size_t)(k + nb) + (size_t)k * M;
            potrf_panel_kernel<<<(rows + 63) / 64, 64, 0, ctx->stream>>>(Akk, P, M, nb, rows, ctx->info_dev);
            // trailing update (lower tiles only): A22 -= P P'
            int rc = gemm(ctx, 0, 1, rows, rows, nb, -1.0, P, M, P, M, 1.0, A + (size_t)(k + nb) + (size_t)(k + nb) * M, M, 1);
            if (rc) return rc;
        }
    }
    zero_upper_kernel<<<(unsigned)(((size_t)M * M + 255) / 256), 256, 0, ctx->stream>>>(A, M);
    int info = 0;
    SGP_CUDA(ctx, cudaMemcpyAsync(&info, ctx->info_dev, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    SGP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (info != 0) {
        char buf[128];
        snprintf(buf, sizeof buf, "Cholesky: non-positive pivot at row %d of %d", info, M);
        SGP_FAIL(ctx, SGP_ERR_NOT_PD, buf);
    }
    return SGP_OK;
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
