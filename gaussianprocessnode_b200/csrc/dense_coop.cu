// M x M factorisations as single cooperative kernels: blocked Cholesky and triangular inverse of the K_uu / posterior
// matrices (M = 20 ... 2048: latency-bound).  One persistent grid, 64 x 64 DMMA tile tasks dealt over the CTAs, grid
// barriers between the dependent phases -- instead of ~6 launches (and a host round trip) per 64-wide panel.
//
//   potrf_coop_kernel   right-looking: [factor diagonal block + its inverse (one CTA, shared memory)] -> panel = A21 Dinv'
//                       (tile GEMM) -> trailing update (lower tiles); the next diagonal block is factorised by the CTA that
//                       updated it, inside the same phase: two grid barriers per panel.
//   trtri_coop_kernel   X = L^-1 by recursive doubling from the diagonal-block inverses the Cholesky already produced:
//                       X21 = -X22 (L21 X11), two GEMM phases per level, log2(M/64) levels; optionally S = X' X (the inverse
//                       of the factorised matrix, lower tiles computed with the triangular K range and mirrored).
// Replaces LAPACK potrf / potri behind fastcholesky! / cholinv (GPnode/UniSGPnode.jl:66-68; regression_kin40k.ipynb:183-184;
// Pendulum_Wishart_2d.ipynb:2542-2543).
#include "sgp_internal.cuh"
#include <cooperative_groups.h>
#include <algorithm>

namespace cg = cooperative_groups;

int sgp_gemm(sgp_ctx* ctx, int opA, int opB, int m, int n, int k, double alpha, const double* A, int lda, const double* B, int ldb,
             double beta, double* C, int ldc, int lower_only);

namespace {

constexpr int TB = 64;        // tile / panel width
constexpr int CT = 256;       // threads per CTA (8 warps: 4 x 2 over the tile, warp tile 16 x 32)
constexpr int KC = 32;        // K chunk staged in shared memory
constexpr int LDA_S = KC + 4; // As[r][k]: r*36 + k   -> conflict-free DMMA A fragments
constexpr int LDB_S = TB + 4; // Bs[k][c]: k*68 + c   -> conflict-free DMMA B fragments
constexpr int SMEM_DOUBLES = TB * (TB + 1) + TB * LDA_S + KC * LDB_S;   // diagonal-block scratch + GEMM staging

struct Acc { double v[2][4][2]; };

__device__ __forceinline__ void acc_zero(Acc& a) {
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) a.v[i][j][0] = a.v[i][j][1] = 0.0;
}

// acc += sum_{k < K} A(r, k) B(k, c) for the 64 x 64 tile; A(r,k) = A[r*a_rs + k*a_ks] (r < rows), B(k,c) = B[k*b_ks + c*b_cs]
// (c < cols); out-of-range rows / columns read as zero.
__device__ void tile_mma(Acc& acc, const double* __restrict__ A, size_t a_rs, size_t a_ks, int rows, const double* __restrict__ B, size_t b_ks,
                         size_t b_cs, int cols, int K, double* __restrict__ As, double* __restrict__ Bs) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wr = warp >> 1, wc = warp & 1;
    for (int k0 = 0; k0 < K; k0 += KC) {
        const int kc = min(KC, K - k0);
        __syncthreads();
        for (int e = tid; e < TB * KC; e += CT) {
            int r, kk;
            if (a_rs == 1) { r = e % TB; kk = e / TB; } else { kk = e % KC; r = e / KC; }
            As[r * LDA_S + kk] = (r < rows && kk < kc) ? A[(size_t)r * a_rs + (size_t)(k0 + kk) * a_ks] : 0.0;
        }
        for (int e = tid; e < TB * KC; e += CT) {
            int c, kk;
            if (b_cs == 1) { c = e % TB; kk = e / TB; } else { kk = e % KC; c = e / KC; }
            Bs[kk * LDB_S + c] = (c < cols && kk < kc) ? B[(size_t)(k0 + kk) * b_ks + (size_t)c * b_cs] : 0.0;
        }
        __syncthreads();
#pragma unroll
        for (int ks = 0; ks < KC / 4; ++ks) {
            double a[2], b[4];
#pragma unroll
            for (int i = 0; i < 2; ++i) a[i] = As[(wr * 16 + 8 * i + (lane >> 2)) * LDA_S + ks * 4 + (lane & 3)];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = Bs[(ks * 4 + (lane & 3)) * LDB_S + wc * 32 + 8 * j + (lane >> 2)];
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma884(acc.v[i][j][0], acc.v[i][j][1], a[i], b[j]);
        }
    }
}

// C(r, c) = alpha * acc + beta * C(r, c) for r < rows, c < cols; C column-major with leading dimension ldc; optionally the
// transposed copy Ct(c, r) = same value (mirror of a symmetric result)
__device__ void tile_store(const Acc& acc, double* __restrict__ C, int ldc, int rows, int cols, double alpha, double beta, double* __restrict__ Ct = nullptr,
                           bool lower_only = false) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int wr = warp >> 1, wc = warp & 1;
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int r = wr * 16 + 8 * i + (lane >> 2), c = wc * 32 + 8 * j + 2 * (lane & 3) + h;
                if (r < rows && c < cols && (!lower_only || c <= r)) {
                    double* p = C + (size_t)r + (size_t)c * ldc;
                    const double v = fma(alpha, acc.v[i][j][h], beta == 0.0 ? 0.0 : beta * *p);
                    *p = v;
                    if (Ct) Ct[(size_t)c + (size_t)r * ldc] = v;
                }
            }
}

// Factor the nb x nb diagonal block at A (leading dimension lda) in shared memory T[64][65], write L back (lower part) and
// its inverse into Dinv (64 x 64 column-major, lower; identity-padded beyond nb).  One CTA.  Two-level: eight 8-wide
// sub-panels, each factorised with a thread per row (registers), followed by a rank-8 update of the trailing block.
__device__ void factor_diag_block(double* __restrict__ A, int lda, int nb, int row0, double* __restrict__ Dinv, int* __restrict__ info, double* __restrict__ T) {
    constexpr int LD = TB + 1, SP = 8;
    __shared__ double rdiag[TB];
    const int tid = threadIdx.x;
    __syncthreads();
    for (int e = tid; e < TB * TB; e += CT) {
        const int r = e % TB, c = e / TB;
        double v = (r == c) ? 1.0 : 0.0;                     // identity padding beyond nb
        if (r < nb && c < nb && c <= r) v = A[(size_t)r + (size_t)c * lda];
        T[r * LD + c] = v;
    }
    __syncthreads();
    for (int jb = 0; jb < TB; jb += SP) {
        // (a) sub-panel (rows jb..63, columns jb..jb+7) by warp 0 alone, warp-synchronously: lane owns rows lane and lane+32,
        //     pivots and multipliers travel by shuffle -- no block barrier inside the 8 pivot steps
        if (tid < 32) {
            const int lane = tid, r0 = lane, r1 = lane + 32;
            const bool hi = jb >= 32;                         // the sub-panel's diagonal rows live in the second half
            double a0[SP], a1[SP];
#pragma unroll
            for (int t = 0; t < SP; ++t) { a0[t] = T[r0 * LD + jb + t]; a1[t] = T[r1 * LD + jb + t]; }
#pragma unroll
            for (int jj = 0; jj < SP; ++jj) {
                const int piv = jb + jj;
                double d = __shfl_sync(0xffffffffu, hi ? a1[jj] : a0[jj], piv & 31);
                if (!(d > 0.0)) { if (lane == 0 && piv < nb) atomicCAS(info, 0, row0 + piv + 1); d = 1.0; }
                const double ri = rsqrt(d), sq = d * ri;
                if (r0 > piv) a0[jj] *= ri; else if (r0 == piv) a0[jj] = sq;
                if (r1 > piv) a1[jj] *= ri; else if (r1 == piv) a1[jj] = sq;
#pragma unroll
                for (int t = jj + 1; t < SP; ++t) {
                    const double v = __shfl_sync(0xffffffffu, hi ? a1[jj] : a0[jj], (jb + t) & 31);     // L[jb+t][piv]
                    if (r0 > piv) a0[t] = fma(-a0[jj], v, a0[t]);
                    if (r1 > piv) a1[t] = fma(-a1[jj], v, a1[t]);
                }
            }
#pragma unroll
            for (int t = 0; t < SP; ++t) {
                if (r0 >= jb) T[r0 * LD + jb + t] = a0[t];
                if (r1 >= jb) T[r1 * LD + jb + t] = a1[t];
            }
        }
        __syncthreads();
        // (b) trailing update: T[i][l] -= sum_t T[i][jb+t] T[l][jb+t] for jb+8 <= l <= i; thread = (row i, column group)
        {
            const int i = tid & (TB - 1), lg = tid >> 6, n0 = jb + SP;
            if (i >= n0) {
                double ri[SP];
#pragma unroll
                for (int t = 0; t < SP; ++t) ri[t] = T[i * LD + jb + t];
                for (int l = n0 + lg; l <= i; l += CT / TB) {
                    double acc = 0.0;
#pragma unroll
                    for (int t = 0; t < SP; ++t) acc = fma(ri[t], T[l * LD + jb + t], acc);
                    T[i * LD + l] -= acc;
                }
            }
        }
        __syncthreads();
    }
    for (int e = tid; e < nb * nb; e += CT) {
        const int r = e % nb, c = e / nb;
        if (c <= r) A[(size_t)r + (size_t)c * lda] = T[r * LD + c];
    }
    if (tid < TB) rdiag[tid] = 1.0 / T[tid * LD + tid];
    __syncthreads();
    // inverse of the unit-padded 64 x 64 lower-triangular block: 4 lanes per column c (lane q owns rows i = q mod 4), axpy
    // form over ALL 64 elimination steps with compile-time indices (steps l < c multiply zeros: no predicates, no divisions)
    {
        const int c = tid >> 2, q = tid & 3;
        const int gbase = (tid & 31) & ~3;
        const unsigned gmask = 0xfu << gbase;
        double x[TB / 4];
#pragma unroll
        for (int s = 0; s < TB / 4; ++s) x[s] = (4 * s + q == c) ? 1.0 : 0.0;
#pragma unroll
        for (int lb = 0; lb < TB / 4; ++lb) {
#pragma unroll
            for (int lq = 0; lq < 4; ++lq) {
                const int l = 4 * lb + lq;
                double xl = x[lb] * rdiag[l];
                if (q == lq) x[lb] = xl;
                xl = __shfl_sync(gmask, xl, gbase | lq);
                if (q > lq) x[lb] = fma(-T[(4 * lb + q) * LD + l], xl, x[lb]);
#pragma unroll
                for (int s = lb + 1; s < TB / 4; ++s) x[s] = fma(-T[(4 * s + q) * LD + l], xl, x[s]);
            }
        }
#pragma unroll
        for (int s = 0; s < TB / 4; ++s) {
            const int i = 4 * s + q;
            Dinv[(size_t)i + (size_t)c * TB] = (i >= c) ? x[s] : 0.0;
        }
    }
    __threadfence();
    __syncthreads();
}

// In-place lower Cholesky of the column-major M x M matrix A (upper triangle zeroed); Dinv[b] = inverse of diagonal block b.
__global__ void __launch_bounds__(CT, 1) potrf_coop_kernel(double* __restrict__ A, int M, double* __restrict__ Dinv, int* __restrict__ info) {
    extern __shared__ double sm[];
    double* T = sm; double* As = sm + TB * (TB + 1); double* Bs = As + TB * LDA_S;
    cg::grid_group grid = cg::this_grid();
    const int nblk = (M + TB - 1) / TB, ncta = gridDim.x, cta = blockIdx.x;
#ifdef SGP_DENSE_CLOCKS
    long long tc[5] = {0, 0, 0, 0, 0}, t0 = clock64(), t1;
#define DCLK(i) do { t1 = clock64(); tc[i] += t1 - t0; t0 = t1; } while (0)
#else
#define DCLK(i) do { } while (0)
#endif
    if (cta == 0) factor_diag_block(A, M, min(TB, M), 0, Dinv, info, T);
    DCLK(0);
    grid.sync();
    DCLK(4);
    for (int k = 0; k < nblk; ++k) {
        const int k0 = k * TB, nb = min(TB, M - k0);
        const double* Dk = Dinv + (size_t)k * TB * TB;
        // panel: L[i, k] = A[i, k] Dk'  for the row blocks below
        for (int i = k + 1 + cta; i < nblk; i += ncta) {
            const int rows = min(TB, M - i * TB);
            double* Aik = A + (size_t)i * TB + (size_t)k0 * M;
            Acc acc; acc_zero(acc);
            tile_mma(acc, Aik, 1, (size_t)M, rows, Dk, (size_t)TB, 1, nb, nb, As, Bs);      // B(kk, c) = Dk(c, kk) = Dk[c + kk*64]
            __syncthreads();
            tile_store(acc, Aik, M, rows, nb, 1.0, 0.0);
        }
        __threadfence();
        DCLK(1);
        grid.sync();
        DCLK(4);
        // trailing update A[i, j] -= L[i, k] L[j, k]' over the lower tiles k < j <= i; tile 0 = (k+1, k+1) goes to CTA 0, which
        // then factorises it (the next panel's diagonal block) before taking its other tiles
        const int nt = nblk - k - 1;
        const int ntiles = nt * (nt + 1) / 2;
        for (int t = cta; t < ntiles; t += ncta) {
            int ii = 0;                                      // tile t -> (ii, jj), jj <= ii, of the trailing block grid
            while ((ii + 1) * (ii + 2) / 2 <= t) ++ii;
            const int jj = t - ii * (ii + 1) / 2;
            const int i = k + 1 + ii, j = k + 1 + jj;
            const int rows = min(TB, M - i * TB), cols = min(TB, M - j * TB);
            Acc acc; acc_zero(acc);
            tile_mma(acc, A + (size_t)i * TB + (size_t)k0 * M, 1, (size_t)M, rows, A + (size_t)j * TB + (size_t)k0 * M, (size_t)M, 1, cols, nb, As, Bs);
            __syncthreads();
            tile_store(acc, A + (size_t)i * TB + (size_t)j * TB * M, M, rows, cols, -1.0, 1.0);
            if (t == 0) {
                __threadfence();
                DCLK(2);
                factor_diag_block(A + (size_t)(k + 1) * TB * ((size_t)M + 1), M, min(TB, M - (k + 1) * TB), (k + 1) * TB,
                                  Dinv + (size_t)(k + 1) * TB * TB, info, T);
                DCLK(0);
            }
        }
        __threadfence();
        DCLK(2);
        grid.sync();
        DCLK(4);
    }
#ifdef SGP_DENSE_CLOCKS
    if (cta == 0 && threadIdx.x == 0) printf("potrf M=%d grid=%d: clocks factor %lld | panel %lld | trailing %lld | grid.sync %lld\n", M, ncta, tc[0], tc[1], tc[2], tc[4]);
#endif
    for (size_t e = (size_t)cta * CT + threadIdx.x; e < (size_t)M * M; e += (size_t)ncta * CT)
        if (e / M > e % M) A[e] = 0.0;
}

// X (lower) = L^-1 from the diagonal-block inverses; Tmp: M x M scratch; S (optional) = X' X = (L L')^-1, full symmetric.
__global__ void __launch_bounds__(CT, 1) trtri_coop_kernel(const double* __restrict__ L, const double* __restrict__ Dinv, double* __restrict__ X,
                                                          double* __restrict__ Tmp, double* __restrict__ S, int M) {
    extern __shared__ double sm[];
    double* As = sm + TB * (TB + 1); double* Bs = As + TB * LDA_S;
    cg::grid_group grid = cg::this_grid();
    const int nblk = (M + TB - 1) / TB, ncta = gridDim.x, cta = blockIdx.x, tid = threadIdx.x;
    // diagonal blocks of X
    for (int b = cta; b < nblk; b += ncta) {
        const int nb = min(TB, M - b * TB);
        for (int e = tid; e < nb * nb; e += CT) {
            const int r = e % nb, c = e / nb;
            X[(size_t)(b * TB + r) + (size_t)(b * TB + c) * M] = Dinv[(size_t)b * TB * TB + r + (size_t)c * TB];
        }
    }
    __threadfence();
    grid.sync();
    for (int bs = 1; bs < nblk; bs *= 2) {                      // merge blocks of bs tiles into blocks of 2*bs tiles
        const int npairs = (nblk + 2 * bs - 1) / (2 * bs);
        // phase A: Tmp21 = L21 X11     (tile (i, j): i in the bottom half, j in the top half; K over the top half, k >= j)
        // phase B: X21 = -X22 Tmp21    (K over the bottom half, k <= i)
        for (int phase = 0; phase < 2; ++phase) {
            int task = 0;
            for (int pr = 0; pr < npairs; ++pr) {
                const int top = pr * 2 * bs, mid = top + bs, bot = min(top + 2 * bs, nblk);
                if (mid >= nblk) continue;
                for (int i = mid; i < bot; ++i)
                    for (int j = top; j < mid; ++j, ++task) {
                        if (task % ncta != cta) continue;
                        const int rows = min(TB, M - i * TB), cols = TB;
                        Acc acc; acc_zero(acc);
                        if (phase == 0) {
                            const int kb = j * TB, K = mid * TB - kb;            // X11(k, j-tile) is zero for k < j-tile
                            tile_mma(acc, L + (size_t)i * TB + (size_t)kb * M, 1, (size_t)M, rows, X + (size_t)kb + (size_t)j * TB * M, 1, (size_t)M, cols, K, As, Bs);
                            __syncthreads();
                            tile_store(acc, Tmp + (size_t)i * TB + (size_t)j * TB * M, M, rows, cols, 1.0, 0.0);
                        } else {
                            const int kb = mid * TB, K = min((i + 1) * TB, M) - kb;  // X22(i-tile, k) is zero for k > i-tile
                            tile_mma(acc, X + (size_t)i * TB + (size_t)kb * M, 1, (size_t)M, rows, Tmp + (size_t)kb + (size_t)j * TB * M, 1, (size_t)M, cols, K, As, Bs);
                            __syncthreads();
                            tile_store(acc, X + (size_t)i * TB + (size_t)j * TB * M, M, rows, cols, -1.0, 0.0);
                        }
                    }
            }
            __threadfence();
            grid.sync();
        }
    }
    if (S) {                                                   // S[i, j] = sum_{k >= i} X[k, i]' X[k, j],  i >= j; mirrored
        const int ntiles = nblk * (nblk + 1) / 2;
        for (int t = cta; t < ntiles; t += ncta) {
            int i = 0;
            while ((i + 1) * (i + 2) / 2 <= t) ++i;
            const int j = t - i * (i + 1) / 2;
            const int rows = min(TB, M - i * TB), cols = min(TB, M - j * TB);
            const int kb = i * TB, K = M - kb;
            Acc acc; acc_zero(acc);
            tile_mma(acc, X + (size_t)kb + (size_t)i * TB * M, (size_t)M, 1, rows, X + (size_t)kb + (size_t)j * TB * M, 1, (size_t)M, cols, K, As, Bs);
            __syncthreads();
            tile_store(acc, S + (size_t)i * TB + (size_t)j * TB * M, M, rows, cols, 1.0, 0.0, S + (size_t)j * TB + (size_t)i * TB * M, i == j);
        }
    }
}

// General GEMM on the same tile routine: C[m x n] = beta C + alpha op(A) op(B), one CTA per 64 x 64 tile (grid-strided).
struct Gemm2Args {
    const double* A; const double* B; double* C;
    int m, n, k, lda, ldb, ldc, opA, opB, lower_only;
    double alpha, beta;
};
__global__ void __launch_bounds__(CT) gemm2_kernel(const Gemm2Args g) {
    extern __shared__ double sm[];
    double* As = sm; double* Bs = sm + TB * LDA_S;
    const int tm = (g.m + TB - 1) / TB, tn = (g.n + TB - 1) / TB;
    for (int t = blockIdx.x; t < tm * tn; t += gridDim.x) {
        const int ti = t % tm, tj = t / tm;
        if (g.lower_only && tj > ti) continue;
        const int rows = min(TB, g.m - ti * TB), cols = min(TB, g.n - tj * TB);
        // A(r, k): opA == 0 -> A[(ti*64 + r) + k*lda], else A[k + (ti*64 + r)*lda];  B(k, c): opB == 0 -> B[k + (tj*64 + c)*ldb], else B[(tj*64 + c) + k*ldb]
        const double* Ab = g.opA == 0 ? g.A + (size_t)ti * TB : g.A + (size_t)ti * TB * g.lda;
        const double* Bb = g.opB == 0 ? g.B + (size_t)tj * TB * g.ldb : g.B + (size_t)tj * TB;
        Acc acc; acc_zero(acc);
        tile_mma(acc, Ab, g.opA == 0 ? 1 : (size_t)g.lda, g.opA == 0 ? (size_t)g.lda : 1, rows, Bb, g.opB == 0 ? 1 : (size_t)g.ldb,
                 g.opB == 0 ? (size_t)g.ldb : 1, cols, g.k, As, Bs);
        __syncthreads();
        tile_store(acc, g.C + (size_t)ti * TB + (size_t)tj * TB * g.ldc, g.ldc, rows, cols, g.alpha, g.beta);
    }
}

int coop_grid(sgp_ctx* ctx, const void* kern, int want) {
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, CT, SMEM_DOUBLES * sizeof(double)) != cudaSuccess || per_sm < 1) per_sm = 1;
    return std::max(1, std::min(want, per_sm * ctx->num_sms));
}

}  // namespace

// In-place lower Cholesky (column-major, upper triangle zeroed); the inverses of the 64 x 64 diagonal blocks are left in
// ctx->dinv_dev for sgp_trtri_lower.  Non-positive pivot -> SGP_ERR_NOT_PD.
int sgp_potrf_lower(sgp_ctx* ctx, double* A, int M) {
    const int nblk = (M + TB - 1) / TB;
    int rc = sgp_ensure(ctx, &ctx->dinv_dev, &ctx->dinv_cap, (size_t)nblk * TB * TB); if (rc) return rc;
    SGP_CUDA(ctx, cudaMemsetAsync(ctx->info_dev, 0, sizeof(int), ctx->stream));
    SGP_CUDA(ctx, cudaFuncSetAttribute(potrf_coop_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(SMEM_DOUBLES * sizeof(double))));
    const int grid = coop_grid(ctx, (const void*)potrf_coop_kernel, std::max(1, (nblk - 1) * nblk / 2));
    double* dinv = ctx->dinv_dev; int* info_dev = ctx->info_dev;
    void* args[] = {&A, &M, &dinv, &info_dev};
    SGP_CUDA(ctx, cudaLaunchCooperativeKernel((const void*)potrf_coop_kernel, dim3(grid), dim3(CT), args, SMEM_DOUBLES * sizeof(double), ctx->stream));
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

int sgp_gemm2(sgp_ctx* ctx, int opA, int opB, int m, int n, int k, double alpha, const double* A, int lda, const double* B, int ldb, double beta,
              double* C, int ldc, int lower_only) {
    if (m <= 0 || n <= 0) return SGP_OK;
    Gemm2Args g{A, B, C, m, n, k, lda, ldb, ldc, opA, opB, lower_only, alpha, beta};
    const int tiles = ((m + TB - 1) / TB) * ((n + TB - 1) / TB);
    const size_t smem = (size_t)(TB * LDA_S + KC * LDB_S) * sizeof(double);
    gemm2_kernel<<<std::min(tiles, 4 * ctx->num_sms), CT, smem, ctx->stream>>>(g);
    SGP_CUDA(ctx, cudaGetLastError());
    return SGP_OK;
}

// B (M x nrhs, ld M) <- L^-1 B (trans = false) or L^-T B (trans = true) with the diagonal-block inverses `dinv` of L
// (as left by sgp_potrf_lower): per 64-wide panel one small GEMM with the inverse block and one GEMM update.
int sgp_trsm_lower_dinv(sgp_ctx* ctx, const double* L, const double* dinv, double* B, double* tmp /* 64 x nrhs */, int M, int nrhs, bool trans) {
    const int nblk = (M + TB - 1) / TB;
    for (int s = 0; s < nblk; ++s) {
        const int kb = trans ? nblk - 1 - s : s;
        const int k = kb * TB, nb = std::min(TB, M - k);
        const double* Dk = dinv + (size_t)kb * TB * TB;
        // X_k = Dk B_k (or Dk' B_k): into tmp, then back
        int rc = sgp_gemm(ctx, trans ? 1 : 0, 0, nb, nrhs, nb, 1.0, Dk, TB, B + k, M, 0.0, tmp, TB, 0); if (rc) return rc;
        SGP_CUDA(ctx, cudaMemcpy2DAsync(B + k, (size_t)M * sizeof(double), tmp, (size_t)TB * sizeof(double), (size_t)nb * sizeof(double), (size_t)nrhs,
                                        cudaMemcpyDeviceToDevice, ctx->stream));
        if (!trans) {
            const int rows = M - k - nb;
            if (rows > 0) { rc = sgp_gemm(ctx, 0, 0, rows, nrhs, nb, -1.0, L + (size_t)(k + nb) + (size_t)k * M, M, B + k, M, 1.0, B + k + nb, M, 0); if (rc) return rc; }
        } else if (k > 0) {
            rc = sgp_gemm(ctx, 1, 0, k, nrhs, nb, -1.0, L + (size_t)k, M, B + k, M, 1.0, B, M, 0); if (rc) return rc;
        }
    }
    return SGP_OK;
}

// X (lower, M x M) = L^-1 for the factor produced by the LAST sgp_potrf_lower call (uses its diagonal-block inverses);
// S (optional) = X' X = (L L')^-1 as a full symmetric matrix.  Tmp: M x M scratch.  X's strict upper triangle is not written.
int sgp_trtri_lower(sgp_ctx* ctx, const double* L, double* X, double* Tmp, double* S, int M) {
    const int nblk = (M + TB - 1) / TB;
    SGP_CUDA(ctx, cudaFuncSetAttribute(trtri_coop_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(SMEM_DOUBLES * sizeof(double))));
    const int grid = coop_grid(ctx, (const void*)trtri_coop_kernel, std::max(1, nblk * (nblk + 1) / 2));
    const double* dinv = ctx->dinv_dev;
    void* args[] = {&L, &dinv, &X, &Tmp, &S, &M};
    SGP_CUDA(ctx, cudaLaunchCooperativeKernel((const void*)trtri_coop_kernel, dim3(grid), dim3(CT), args, SMEM_DOUBLES * sizeof(double), ctx->stream));
    SGP_CUDA(ctx, cudaGetLastError());
    return SGP_OK;
}
