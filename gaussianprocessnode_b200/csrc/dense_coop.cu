// M x M factorisations of the path as ONE cooperative kernel per job (M = 20 ... 2048: latency-bound, so the design minimises the
// critical path, not the FLOP count):
//
//   dense_job_kernel   [build A] -> blocked right-looking Cholesky -> [X = L^-1 by recursive doubling, S = X'X] -> [mu = S xi]
//                      -> [transposed copy of L]
//       build:  A = Lambda_prior + w Psi2 (the N-th `prod`), A = K_uu(Z) + jitter I, A = Sigma + mu mu', or A as given
//       Cholesky, per 64-wide panel:  panel L21 = A21 Dinv'  (16 x 64 row strips, one per CTA)  | grid barrier |
//                      trailing update in 32 x 32 sub-tiles over all CTAs, while CTA 0 updates the next diagonal block straight into
//                      shared memory and factorises it there | grid barrier
//       diagonal block (64 x 64, one CTA, shared memory): two 32 x 32 Cholesky factorisations by ONE WARP with the block's rows in
//                      registers (pivots and multipliers travel by shuffle: ~110 clocks per pivot, no block barrier inside), their
//                      inverses by one warp (a lane per column), and four 32^3 DMMA products for the off-diagonal parts
//       every GEMM-shaped piece is a `gemm_task`: global -> registers -> shared-memory staging (software pipelined), DMMA.8x8x4
//
// Replaces LAPACK potrf / potri behind fastcholesky! / cholinv and the posterior update of the N-th `prod`
// (GPnode/UniSGPnode.jl:62-73; regression_kin40k.ipynb:183-184; Pendulum_Wishart_2d.ipynb:2542-2543).
#include "sgp_internal.cuh"
#include <cooperative_groups.h>
#include <algorithm>
#include <cstdlib>

namespace cg = cooperative_groups;

int sgp_gemm(sgp_ctx* ctx, int opA, int opB, int m, int n, int k, double alpha, const double* A, int lda, const double* B, int ldb,
             double beta, double* C, int ldc, int lower_only);

namespace {

constexpr int TB = 64;        // panel width = diagonal block
constexpr int CT = 256;       // threads per CTA (8 warps)
constexpr int KC = 32;        // K chunk staged in shared memory
constexpr int LDT = TB + 4;   // leading dimension of the 64 x 64 shared-memory blocks (= 4 mod 16: conflict-free DMMA fragment loads)
constexpr int STAGE_A = TB * (KC + 4), STAGE_B = KC * (TB + 4);
constexpr int SMEM_DOUBLES = 2 * TB * LDT + TB + STAGE_A + STAGE_B;     // T | Xi | rdiag | As | Bs

// ---- tile GEMM task: acc (+)= sum_{k < K} A(r, k) B(k, c) for a TR x TC tile, 8 warps arranged WR x (8 / WR) ------------------------
template <int TR, int TC, int WR>
struct Tile {
    static constexpr int WC = 8 / WR, WTR = TR / WR, WTC = TC / WC, MI = WTR / 8, NJ = WTC / 8;
    static constexpr int LDA = KC + 4, LDB = TC + 4;              // both = 4 (mod 16)
    static constexpr int AP = TR * KC / CT, BP = KC * TC / CT;   // staged elements per thread
    static_assert(WTR % 8 == 0 && WTC % 8 == 0 && AP >= 1 && BP >= 1, "tile shape");
    struct Acc { double v[MI][NJ][2]; };
};

template <class TL>
__device__ __forceinline__ void acc_zero(typename TL::Acc& a) {
#pragma unroll
    for (int i = 0; i < TL::MI; ++i)
#pragma unroll
        for (int j = 0; j < TL::NJ; ++j) a.v[i][j][0] = a.v[i][j][1] = 0.0;
}

// A(r, k) = A[r * a_rs + k * a_ks] (r < rows), B(k, c) = B[k * b_ks + c * b_cs] (c < cols); out-of-range elements read as zero.
// Global loads of chunk k+1 are in flight while chunk k is multiplied.  All 256 threads must call it.
template <class TL>
__device__ __forceinline__ void gemm_task(typename TL::Acc& acc, const double* __restrict__ A, size_t a_rs, size_t a_ks, int rows,
                                          const double* __restrict__ B, size_t b_ks, size_t b_cs, int cols, int K, double* __restrict__ As,
                                          double* __restrict__ Bs) {
    constexpr int TR = TL::WTR * (8 / TL::WC), TC = TL::WTC * TL::WC;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wr = warp / TL::WC, wc = warp % TL::WC;
    double ra[TL::AP], rb[TL::BP];
    const bool a_rowfast = a_rs == 1, b_colfast = b_cs == 1;
    auto gload = [&](int k0) {
#pragma unroll
        for (int q = 0; q < TL::AP; ++q) {
            const int e = tid + q * CT;
            const int r = a_rowfast ? e % TR : e / KC, kk = a_rowfast ? e / TR : e % KC;
            ra[q] = (r < rows && k0 + kk < K) ? A[(size_t)r * a_rs + (size_t)(k0 + kk) * a_ks] : 0.0;
        }
#pragma unroll
        for (int q = 0; q < TL::BP; ++q) {
            const int e = tid + q * CT;
            const int c = b_colfast ? e % TC : e / KC, kk = b_colfast ? e / TC : e % KC;
            rb[q] = (c < cols && k0 + kk < K) ? B[(size_t)(k0 + kk) * b_ks + (size_t)c * b_cs] : 0.0;
        }
    };
    auto sstore = [&]() {
#pragma unroll
        for (int q = 0; q < TL::AP; ++q) {
            const int e = tid + q * CT;
            const int r = a_rowfast ? e % TR : e / KC, kk = a_rowfast ? e / TR : e % KC;
            As[r * TL::LDA + kk] = ra[q];
        }
#pragma unroll
        for (int q = 0; q < TL::BP; ++q) {
            const int e = tid + q * CT;
            const int c = b_colfast ? e % TC : e / KC, kk = b_colfast ? e / TC : e % KC;
            Bs[kk * TL::LDB + c] = rb[q];
        }
    };
    gload(0);
    for (int k0 = 0; k0 < K; k0 += KC) {
        __syncthreads();                       // the previous chunk's fragments have been read
        sstore();
        __syncthreads();
        if (k0 + KC < K) gload(k0 + KC);
#pragma unroll
        for (int ks = 0; ks < KC / 4; ++ks) {
            double a[TL::MI], b[TL::NJ];
#pragma unroll
            for (int i = 0; i < TL::MI; ++i) a[i] = As[(wr * TL::WTR + 8 * i + (lane >> 2)) * TL::LDA + ks * 4 + (lane & 3)];
#pragma unroll
            for (int j = 0; j < TL::NJ; ++j) b[j] = Bs[(ks * 4 + (lane & 3)) * TL::LDB + wc * TL::WTC + 8 * j + (lane >> 2)];
#pragma unroll
            for (int i = 0; i < TL::MI; ++i)
#pragma unroll
                for (int j = 0; j < TL::NJ; ++j) dmma884(acc.v[i][j][0], acc.v[i][j][1], a[i], b[j]);
        }
    }
}

// C(r, c) = alpha acc + beta C(r, c), r < rows, c < cols; C column-major (ldc).  lower_only: only elements with grow0 + r >= gcol0 + c.
// Ct (optional): the same value to Ct(c, r) (mirror of a symmetric result).
template <class TL>
__device__ __forceinline__ void store_task(const typename TL::Acc& acc, double* __restrict__ C, int ldc, int rows, int cols, double alpha,
                                           double beta, double* __restrict__ Ct = nullptr, bool lower_only = false, int grow0 = 0,
                                           int gcol0 = 0) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int wr = warp / TL::WC, wc = warp % TL::WC;
#pragma unroll
    for (int i = 0; i < TL::MI; ++i)
#pragma unroll
        for (int j = 0; j < TL::NJ; ++j)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int r = wr * TL::WTR + 8 * i + (lane >> 2), c = wc * TL::WTC + 8 * j + 2 * (lane & 3) + h;
                if (r < rows && c < cols && (!lower_only || grow0 + r >= gcol0 + c)) {
                    double* p = C + (size_t)r + (size_t)c * ldc;
                    const double v = fma(alpha, acc.v[i][j][h], beta == 0.0 ? 0.0 : beta * *p);
                    *p = v;
                    if (Ct) Ct[(size_t)c + (size_t)r * ldc] = v;
                }
            }
}

using T64 = Tile<64, 64, 4>;      // warp tile 16 x 32
using T32 = Tile<32, 32, 4>;      // warp tile  8 x 16
using T16 = Tile<16, 64, 2>;      // row strip: warp tile 8 x 16

// ---- 32 x 32 pieces of the diagonal-block factorisation (shared memory, leading dimension LDT) -------------------------------------
// In-place lower Cholesky of the 32 x 32 block at T[(off + i) * LDT + off + c] by ONE warp: lane i holds row i in registers; per pivot
// one shuffle broadcasts the pivot, every lane scales its element, and the multipliers travel by shuffle for the rank-1 update.
// rdiag[off + j] = 1 / L_jj.  A non-positive pivot is recorded (rows below `nvalid` only) and replaced by 1.
__device__ __forceinline__ void chol32_warp(double* __restrict__ T, int off, double* __restrict__ rdiag, int* __restrict__ info, int row0, int nvalid) {
    const int lane = threadIdx.x & 31;
    double a[32];
    double* row = T + (off + lane) * LDT + off;
#pragma unroll
    for (int c = 0; c < 32; ++c) a[c] = row[c];
    double rmine = 1.0;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        double d = __shfl_sync(0xffffffffu, a[j], j);
        if (!(d > 0.0)) { if (lane == 0 && off + j < nvalid) atomicCAS(info, 0, row0 + off + j + 1); d = 1.0; }
        const double r = rsqrt(d);
        double l = a[j] * r;
        if (lane == j) { l = d * r; rmine = r; }
        a[j] = l;
#pragma unroll
        for (int t = j + 1; t < 32; ++t) {
            const double v = __shfl_sync(0xffffffffu, l, t);      // L[t][j]
            a[t] = fma(-l, v, a[t]);
        }
    }
#pragma unroll
    for (int c = 0; c < 32; ++c) row[c] = (c <= lane) ? a[c] : 0.0;
    rdiag[off + lane] = rmine;
}

// Xi block = inverse of the lower-triangular 32 x 32 block of T at `off`, by ONE warp: lane c solves L x = e_c (right-looking, the
// elements of L are broadcast reads).
__device__ __forceinline__ void inv32_warp(const double* __restrict__ T, int off, const double* __restrict__ rdiag, double* __restrict__ Xi) {
    const int lane = threadIdx.x & 31;
    double x[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) x[i] = (i == lane) ? 1.0 : 0.0;
#pragma unroll
    for (int l = 0; l < 32; ++l) {
        x[l] *= rdiag[off + l];
#pragma unroll
        for (int i = l + 1; i < 32; ++i) x[i] = fma(-T[(off + i) * LDT + off + l], x[l], x[i]);
    }
#pragma unroll
    for (int i = 0; i < 32; ++i) Xi[(off + i) * LDT + off + lane] = (i >= lane) ? x[i] : 0.0;
}

// 32 x 32 x 32 product on shared-memory operands by all 8 warps (warp tile 8 x 16): returns the warp's fragment of
//   sum_k A(r, k) B(k, c),  A(r, k) = A[r * a_rs + k * a_ks],  B(k, c) = B[k * b_ks + c * b_cs]
struct Frag32 { double v[2][2]; };
__device__ __forceinline__ Frag32 smem_mma32(const double* __restrict__ A, int a_rs, int a_ks, const double* __restrict__ B, int b_ks, int b_cs) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int wr = warp >> 1, wc = warp & 1;
    Frag32 f;
    f.v[0][0] = f.v[0][1] = f.v[1][0] = f.v[1][1] = 0.0;
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) {
        const int k = ks * 4 + (lane & 3);
        const double a = A[(wr * 8 + (lane >> 2)) * a_rs + k * a_ks];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const double b = B[k * b_ks + (wc * 16 + 8 * j + (lane >> 2)) * b_cs];
            dmma884(f.v[j][0], f.v[j][1], a, b);
        }
    }
    return f;
}
// C(r, c) = alpha f + beta C(r, c) for the warp's fragment; C[r * LDT + c]
__device__ __forceinline__ void smem_store32(const Frag32& f, double* __restrict__ C, double alpha, double beta) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int wr = warp >> 1, wc = warp & 1;
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            double* p = C + (wr * 8 + (lane >> 2)) * LDT + wc * 16 + 8 * j + 2 * (lane & 3) + h;
            *p = fma(alpha, f.v[j][h], beta == 0.0 ? 0.0 : beta * *p);
        }
}

// Factor the nb x nb diagonal block held in T (lower part valid, identity-padded beyond nb, zeros above the diagonal), write L to A
// (global, lower part) and its inverse to Dinv (64 x 64 column-major, identity-padded).  One CTA, all 256 threads.
__device__ void factor_diag_smem(double* __restrict__ T, double* __restrict__ Xi, double* __restrict__ rdiag, double* __restrict__ A, int lda, int nb,
                                 int row0, double* __restrict__ Dinv, int* __restrict__ info) {
    const int tid = threadIdx.x, warp = tid >> 5;
    double* T21 = T + 32 * LDT; double* T22 = T + 32 * LDT + 32;
    double* X21 = Xi + 32 * LDT;
    __syncthreads();
    if (warp == 0) { chol32_warp(T, 0, rdiag, info, row0, nb); __syncwarp(); inv32_warp(T, 0, rdiag, Xi); }
    __syncthreads();
    {   // L21 = A21 X11'  (in place: all reads before the writes)
        const Frag32 f = smem_mma32(T21, LDT, 1, Xi, 1, LDT);
        __syncthreads();
        smem_store32(f, T21, 1.0, 0.0);
    }
    __syncthreads();
    {   // A22 -= L21 L21' ;  W = L21 X11 (kept in the X21 slot)
        const Frag32 f = smem_mma32(T21, LDT, 1, T21, 1, LDT);
        const Frag32 g = smem_mma32(T21, LDT, 1, Xi, LDT, 1);
        smem_store32(f, T22, -1.0, 1.0);
        smem_store32(g, X21, 1.0, 0.0);
    }
    __syncthreads();
    if (warp == 0) { chol32_warp(T, 32, rdiag, info, row0, nb); __syncwarp(); inv32_warp(T, 32, rdiag, Xi); }
    __syncthreads();
    {   // X21 = -X22 W  (in place)
        const Frag32 f = smem_mma32(Xi + 32 * LDT + 32, LDT, 1, X21, LDT, 1);
        __syncthreads();
        smem_store32(f, X21, -1.0, 0.0);
    }
    __syncthreads();
    for (int e = tid; e < nb * nb; e += CT) {
        const int r = e % nb, c = e / nb;
        if (c <= r) A[(size_t)r + (size_t)c * lda] = T[r * LDT + c];
    }
    for (int e = tid; e < TB * TB; e += CT) {
        const int r = e % TB, c = e / TB;
        Dinv[(size_t)r + (size_t)c * TB] = (c <= r) ? Xi[r * LDT + c] : 0.0;
    }
    __threadfence();
    __syncthreads();
}

// T <- the nb x nb block at A (lower part), identity-padded to 64 x 64, zeros above the diagonal
__device__ __forceinline__ void load_diag_smem(double* __restrict__ T, const double* __restrict__ A, int lda, int nb) {
    for (int e = threadIdx.x; e < TB * TB; e += CT) {
        const int r = e % TB, c = e / TB;
        double v = (r == c) ? 1.0 : 0.0;
        if (r < nb && c < nb) v = (c <= r) ? A[(size_t)r + (size_t)c * lda] : 0.0;
        T[r * LDT + c] = v;
    }
}

struct DenseJob {
    int M;
    int build;                 // 0: A as given | 1: A = P + w S2, xi = xip + w s1 (carry: P <- A, xip <- xi) | 2: A = K_uu(Z) + jitter I | 3: A = Sig + mu mu'
    double* A;                 // M x M column-major, factored in place: L in the lower triangle, strict upper triangle zeroed
    double* Dinv;              // inverses of the 64 x 64 diagonal blocks of L
    int* info;                 // 0, or 1 + the row of the first non-positive pivot
    const double* S2; const double* s1; double* P; double* xip; double* xi; double w; int carry;     // build 1
    const double* Z; int D, kind; double variance, jitter; double ell_inv[SGP_MAX_D];               // build 2
    const double* Sig; const double* mu_in;                                                         // build 3
    double* X; double* Tmp; double* S;     // optional: X = L^-1 (lower), S = X' X = (L L')^-1 full symmetric; Tmp = M x M scratch
    double* mu;                            // optional (needs S and xi): mu = S xi
    double* Ut;                            // optional: Ut = L' (upper triangular, strict lower part zero)
    long long* clk;                        // optional: CTA 0's clocks {build, factor, panel, trailing, barriers, inverse, S, tail}
};

__global__ void __launch_bounds__(CT, 1) dense_job_kernel(const __grid_constant__ DenseJob j) {
    extern __shared__ double sm[];
    double* T = sm; double* Xi = T + TB * LDT; double* rdiag = Xi + TB * LDT; double* As = rdiag + TB; double* Bs = As + STAGE_A;
    cg::grid_group grid = cg::this_grid();
    const int M = j.M, nblk = (M + TB - 1) / TB, ncta = gridDim.x, cta = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const size_t MM = (size_t)M * M;
    double* __restrict__ A = j.A;
    long long tc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, t0 = clock64(), t1;
#define DCLK(i) do { if (j.clk) { t1 = clock64(); tc[i] += t1 - t0; t0 = t1; } } while (0)

    // ---- build -------------------------------------------------------------------------------------------------------------------
    if (j.build == 1) {
        for (size_t e = (size_t)cta * CT + tid; e < MM; e += (size_t)ncta * CT) {
            const double a = fma(j.w, j.S2[e], j.P[e]);
            A[e] = a;
            if (j.carry) j.P[e] = a;
        }
        if (cta == ncta - 1)
            for (int e = tid; e < M; e += CT) {
                const double x = fma(j.w, j.s1[e], j.xip[e]);
                j.xi[e] = x;
                if (j.carry) j.xip[e] = x;
            }
    } else if (j.build == 2) {
        for (size_t e = (size_t)cta * CT + tid; e < MM; e += (size_t)ncta * CT) {
            const int r = (int)(e % M), c = (int)(e / M);
            double r2 = 0.0;
            for (int d = 0; d < j.D; ++d) { const double t = (j.Z[(size_t)r * j.D + d] - j.Z[(size_t)c * j.D + d]) * j.ell_inv[d]; r2 = fma(t, t, r2); }
            double v;
            if (j.kind == SGP_KERNEL_SE) v = j.variance * exp(-0.5 * r2);
            else if (j.kind == SGP_KERNEL_MATERN32) { const double s = sqrt(3.0 * r2); v = j.variance * (1.0 + s) * exp(-s); }
            else { const double s = sqrt(5.0 * r2); v = j.variance * (1.0 + s + s * s / 3.0) * exp(-s); }
            if (r == c) v += j.jitter;
            A[e] = v;
        }
    } else if (j.build == 3) {
        for (size_t e = (size_t)cta * CT + tid; e < MM; e += (size_t)ncta * CT) A[e] = fma(j.mu_in[e % M], j.mu_in[e / M], j.Sig[e]);
    }
    if (j.build) { __threadfence(); grid.sync(); }
    DCLK(0);

    // ---- Cholesky ----------------------------------------------------------------------------------------------------------------
    if (cta == 0) {
        load_diag_smem(T, A, M, min(TB, M));
        factor_diag_smem(T, Xi, rdiag, A, M, min(TB, M), 0, j.Dinv, j.info);
    }
    DCLK(1);
    grid.sync();
    DCLK(4);
    const int nwork = ncta > 1 ? ncta - 1 : 1, wid = ncta > 1 ? cta - 1 : 0;      // CTAs that take the trailing sub-tiles (CTA 0 factorises)
    for (int k = 0; k + 1 < nblk; ++k) {
        const int k0 = k * TB, R0 = k0 + TB;                   // trailing matrix starts at row / column R0
        const double* Dk = j.Dinv + (size_t)k * TB * TB;
        // panel: 16-row strips of A[R0:, k0:k0+64] <- strip * Dk'   (in place: a strip is private to its CTA)
        const int nstrips = (M - R0 + 15) / 16;
        for (int s = cta; s < nstrips; s += ncta) {
            const int r0 = R0 + 16 * s, rows = min(16, M - r0);
            double* Ar = A + (size_t)r0 + (size_t)k0 * M;
            T16::Acc acc; acc_zero<T16>(acc);
            gemm_task<T16>(acc, Ar, 1, (size_t)M, rows, Dk, (size_t)TB, 1, TB, TB, As, Bs);      // B(kk, c) = Dk(c, kk)
            store_task<T16>(acc, Ar, M, rows, TB, 1.0, 0.0);
        }
        __threadfence();
        DCLK(2);
        grid.sync();
        DCLK(4);
        // trailing update A[i, c] -= sum_kk L[i, k0 + kk] L[c, k0 + kk] on the lower triangle of A[R0:, R0:]
        const int nbn = min(TB, M - R0);
        if (cta == 0) {     // next diagonal block: updated straight into shared memory and factorised there
            T64::Acc acc; acc_zero<T64>(acc);
            gemm_task<T64>(acc, A + (size_t)R0 + (size_t)k0 * M, 1, (size_t)M, nbn, A + (size_t)R0 + (size_t)k0 * M, (size_t)M, 1, nbn, TB, As, Bs);
            load_diag_smem(T, A + (size_t)R0 * ((size_t)M + 1), M, nbn);
            __syncthreads();
            {
                const int wr = warp / T64::WC, wc = warp % T64::WC;
#pragma unroll
                for (int i = 0; i < T64::MI; ++i)
#pragma unroll
                    for (int jj = 0; jj < T64::NJ; ++jj)
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            const int r = wr * T64::WTR + 8 * i + (lane >> 2), c = wc * T64::WTC + 8 * jj + 2 * (lane & 3) + h;
                            if (r < nbn && c <= r) T[r * LDT + c] -= acc.v[i][jj][h];
                        }
            }
            DCLK(3);
            factor_diag_smem(T, Xi, rdiag, A + (size_t)R0 * ((size_t)M + 1), M, nbn, R0, j.Dinv + (size_t)(k + 1) * TB * TB, j.info);
            DCLK(1);
        }
        if (cta > 0 || ncta == 1) {
            const int n32 = (M - R0 + 31) / 32;
            const int ntask = n32 * (n32 + 1) / 2 - (n32 >= 2 ? 3 : 1);            // the sub-tiles of the first 64 x 64 block belong to CTA 0
            for (int t = wid; t < ntask; t += nwork) {
                const int tt = t + 3;
                int a = 0;
                while ((a + 1) * (a + 2) / 2 <= tt) ++a;
                const int b = tt - a * (a + 1) / 2;
                const int r0 = R0 + 32 * a, c0 = R0 + 32 * b, rows = min(32, M - r0), cols = min(32, M - c0);
                T32::Acc acc; acc_zero<T32>(acc);
                gemm_task<T32>(acc, A + (size_t)r0 + (size_t)k0 * M, 1, (size_t)M, rows, A + (size_t)c0 + (size_t)k0 * M, (size_t)M, 1, cols, TB, As, Bs);
                store_task<T32>(acc, A + (size_t)r0 + (size_t)c0 * M, M, rows, cols, -1.0, 1.0, nullptr, a == b, r0, c0);
            }
        }
        __threadfence();
        DCLK(3);
        grid.sync();
        DCLK(4);
    }

    // ---- X = L^-1 (recursive doubling from the diagonal-block inverses), S = X' X ----------------------------------------------------
    if (j.X) {
        double* __restrict__ X = j.X;
        for (int b = cta; b < nblk; b += ncta) {
            const int nb = min(TB, M - b * TB);
            for (int e = tid; e < nb * nb; e += CT) {
                const int r = e % nb, c = e / nb;
                X[(size_t)(b * TB + r) + (size_t)(b * TB + c) * M] = j.Dinv[(size_t)b * TB * TB + r + (size_t)c * TB];
            }
        }
        __threadfence();
        grid.sync();
        for (int bs = 1; bs < nblk; bs *= 2) {                      // merge blocks of bs tiles into blocks of 2 bs tiles
            const int npairs = (nblk + 2 * bs - 1) / (2 * bs);
            // phase 0: Tmp21 = L21 X11  (K from the column sub-block's first row to the end of the top half)
            // phase 1: X21 = -X22 Tmp21 (K from the start of the bottom half to the row sub-block's last row)
            for (int phase = 0; phase < 2; ++phase) {
                int task = 0;
                for (int pr = 0; pr < npairs; ++pr) {
                    const int top = pr * 2 * bs * TB, mid = top + bs * TB, bot = min(top + 2 * bs * TB, M);
                    if (mid >= M) continue;
                    const int nr = (bot - mid + 31) / 32, ncol = (mid - top) / 32;
                    const int first = ((cta - task) % ncta + ncta) % ncta;         // this CTA's first sub-tile of the pair
                    for (int q = first; q < nr * ncol; q += ncta) {
                        const int r0 = mid + 32 * (q % nr), c0 = top + 32 * (q / nr);
                        const int rows = min(32, M - r0);
                        T32::Acc acc; acc_zero<T32>(acc);
                        if (phase == 0) {
                            gemm_task<T32>(acc, A + (size_t)r0 + (size_t)c0 * M, 1, (size_t)M, rows, X + (size_t)c0 + (size_t)c0 * M, 1, (size_t)M, 32, mid - c0, As, Bs);
                            store_task<T32>(acc, j.Tmp + (size_t)r0 + (size_t)c0 * M, M, rows, 32, 1.0, 0.0);
                        } else {
                            const int K = min(r0 + 32, M) - mid;
                            gemm_task<T32>(acc, X + (size_t)r0 + (size_t)mid * M, 1, (size_t)M, rows, j.Tmp + (size_t)mid + (size_t)c0 * M, 1, (size_t)M, 32, K, As, Bs);
                            store_task<T32>(acc, X + (size_t)r0 + (size_t)c0 * M, M, rows, 32, -1.0, 0.0);
                        }
                    }
                    task += nr * ncol;
                }
                __threadfence();
                grid.sync();
            }
        }
        DCLK(5);
        if (j.S) {        // S[i, c] = sum_{k >= i} X[k, i] X[k, c] for c <= i (32 x 32 sub-tiles, longest K first), mirrored
            const int n32 = (M + 31) / 32, ntask = n32 * (n32 + 1) / 2;
            for (int t = cta; t < ntask; t += ncta) {
                int a = 0;
                while ((a + 1) * (a + 2) / 2 <= t) ++a;
                const int b = t - a * (a + 1) / 2;
                const int r0 = 32 * a, c0 = 32 * b, rows = min(32, M - r0), cols = min(32, M - c0);
                T32::Acc acc; acc_zero<T32>(acc);
                gemm_task<T32>(acc, X + (size_t)r0 + (size_t)r0 * M, (size_t)M, 1, rows, X + (size_t)r0 + (size_t)c0 * M, 1, (size_t)M, cols, M - r0, As, Bs);
                store_task<T32>(acc, j.S + (size_t)r0 + (size_t)c0 * M, M, rows, cols, 1.0, 0.0, j.S + (size_t)c0 + (size_t)r0 * M, a == b, r0, c0);
            }
            if (j.mu) {
                __threadfence();
                grid.sync();
                // mu = S xi: a warp per column of the symmetric S (fixed-shape tree: deterministic)
                for (int i = cta * (CT / 32) + warp; i < M; i += ncta * (CT / 32)) {
                    const double* col = j.S + (size_t)i * M;
                    double v = 0.0;
                    for (int r = lane; r < M; r += 32) v = fma(col[r], j.xi[r], v);
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                    if (lane == 0) j.mu[i] = v;
                }
            }
        }
        DCLK(6);
    }

    // ---- tail: zero the strict upper triangle of L; Ut = L' -----------------------------------------------------------------------------
    if (j.Ut) {
        const int n32 = (M + 31) / 32;
        double* tile = sm;                                         // 32 x 33
        for (int t = cta; t < n32 * n32; t += ncta) {
            const int bi = t % n32, bj = t / n32;                  // block (bi, bj) of L -> block (bj, bi) of Ut
            __syncthreads();
            for (int e = tid; e < 32 * 32; e += CT) {
                const int r = bi * 32 + e % 32, c = bj * 32 + e / 32;
                double v = 0.0;
                if (r < M && c < M && c <= r) v = A[(size_t)r + (size_t)c * M];
                tile[(e % 32) * 33 + e / 32] = v;
            }
            __syncthreads();
            for (int e = tid; e < 32 * 32; e += CT) {
                const int c = bj * 32 + e % 32, r = bi * 32 + e / 32;          // Ut[c, r] = L[r, c]
                if (r < M && c < M) {
                    j.Ut[(size_t)c + (size_t)r * M] = tile[(e / 32) * 33 + e % 32];
                    if (c > r) A[(size_t)r + (size_t)c * M] = 0.0;
                }
            }
        }
    } else {
        for (size_t e = (size_t)cta * CT + tid; e < MM; e += (size_t)ncta * CT)
            if (e / M > e % M) A[e] = 0.0;
    }
    DCLK(7);
    if (j.clk && cta == 0 && tid == 0)
        for (int i = 0; i < 8; ++i) j.clk[i] = tc[i];
#undef DCLK
}

// General GEMM on the same tile routine: C[m x n] = beta C + alpha op(A) op(B); 32 x 32 tiles while they fit one wave, else 64 x 64.
struct Gemm2Args {
    const double* A; const double* B; double* C;
    int m, n, k, lda, ldb, ldc, opA, opB, lower_only;
    double alpha, beta;
};
template <class TL, int TS>
__global__ void __launch_bounds__(CT) gemm2_kernel(const Gemm2Args g) {
    extern __shared__ double sm[];
    double* As = sm; double* Bs = sm + STAGE_A;
    const int tm = (g.m + TS - 1) / TS, tn = (g.n + TS - 1) / TS;
    for (int t = blockIdx.x; t < tm * tn; t += gridDim.x) {
        const int ti = t % tm, tj = t / tm;
        if (g.lower_only && tj > ti) continue;
        const int rows = min(TS, g.m - ti * TS), cols = min(TS, g.n - tj * TS);
        // A(r, k): opA == 0 -> A[(ti*TS + r) + k*lda], else A[k + (ti*TS + r)*lda];  B(k, c): opB == 0 -> B[k + (tj*TS + c)*ldb], else B[(tj*TS + c) + k*ldb]
        const double* Ab = g.opA == 0 ? g.A + (size_t)ti * TS : g.A + (size_t)ti * TS * g.lda;
        const double* Bb = g.opB == 0 ? g.B + (size_t)tj * TS * g.ldb : g.B + (size_t)tj * TS;
        typename TL::Acc acc; acc_zero<TL>(acc);
        gemm_task<TL>(acc, Ab, g.opA == 0 ? 1 : (size_t)g.lda, g.opA == 0 ? (size_t)g.lda : 1, rows, Bb, g.opB == 0 ? 1 : (size_t)g.ldb,
                      g.opB == 0 ? (size_t)g.ldb : 1, cols, g.k, As, Bs);
        store_task<TL>(acc, g.C + (size_t)ti * TS + (size_t)tj * TS * g.ldc, g.ldc, rows, cols, g.alpha, g.beta);
    }
}

}  // namespace

// One cooperative launch of dense_job_kernel on the ctx stream.  The caller checks *info_dev (sgp_dense_info) when it next synchronises.
int sgp_dense_job(sgp_ctx* ctx, const SgpDenseJob& in) {
    DenseJob j{};
    j.M = in.M; j.build = in.build; j.A = in.A; j.Dinv = in.Dinv; j.info = ctx->info_dev;
    j.S2 = in.S2; j.s1 = in.s1; j.P = in.P; j.xip = in.xip; j.xi = in.xi; j.w = in.w; j.carry = in.carry;
    j.Z = ctx->Z_dev; j.D = ctx->D; j.kind = ctx->kind; j.variance = ctx->variance; j.jitter = in.jitter;
    for (int d = 0; d < SGP_MAX_D; ++d) j.ell_inv[d] = d < ctx->D ? 1.0 / ctx->ell[d] : 0.0;
    j.Sig = in.Sig; j.mu_in = in.mu_in; j.X = in.X; j.Tmp = in.Tmp; j.S = in.S; j.mu = in.mu; j.Ut = in.Ut; j.clk = in.clk;
    if (j.mu && !(j.S && j.xi)) SGP_FAIL(ctx, SGP_ERR_ARG, "dense job: mu needs S and xi");
    const int M = in.M, nblk = (M + TB - 1) / TB, n32 = (M + 31) / 32;
    const size_t smem = SMEM_DOUBLES * sizeof(double);
    SGP_CUDA(ctx, cudaFuncSetAttribute(dense_job_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    static const bool print_clocks = std::getenv("SGP_DENSE_CLOCKS") != nullptr;       // tuning aid: per-phase clocks of CTA 0 on stdout
    long long* clk_dev = nullptr;
    if (print_clocks && !j.clk) { SGP_CUDA(ctx, cudaMalloc((void**)&clk_dev, 8 * sizeof(long long))); j.clk = clk_dev; }
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, dense_job_kernel, CT, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
    int want = std::max(1, n32 * (n32 + 1) / 2);                    // the widest phases: trailing sub-tiles of the first panel, S = X'X
    if (nblk == 1 && !in.X) want = 1;
    const int grid = std::max(1, std::min(want, per_sm * ctx->num_sms));
    if (in.reset_info) SGP_CUDA(ctx, cudaMemsetAsync(ctx->info_dev, 0, sizeof(int), ctx->stream));
    void* args[] = {&j};
    SGP_CUDA(ctx, cudaLaunchCooperativeKernel((const void*)dense_job_kernel, dim3(grid), dim3(CT), args, smem, ctx->stream));
    SGP_CUDA(ctx, cudaGetLastError());
    if (clk_dev) {
        long long c[8];
        SGP_CUDA(ctx, cudaMemcpyAsync(c, clk_dev, sizeof c, cudaMemcpyDeviceToHost, ctx->stream));
        SGP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        cudaFree(clk_dev);
        printf("dense job M=%d build=%d grid=%d: clocks build %lld | factor %lld | panel %lld | trailing %lld | barriers %lld | inverse %lld | S+mu %lld | tail %lld\n",
               M, in.build, grid, c[0], c[1], c[2], c[3], c[4], c[5], c[6], c[7]);
        fflush(stdout);
    }
    return SGP_OK;
}

// Synchronises the stream and turns a recorded non-positive pivot into SGP_ERR_NOT_PD.
int sgp_dense_info(sgp_ctx* ctx, const char* what) {
    int info = 0;
    SGP_CUDA(ctx, cudaMemcpyAsync(&info, ctx->info_dev, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    SGP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (info != 0) {
        char buf[160];
        snprintf(buf, sizeof buf, "%s: Cholesky met a non-positive pivot at row %d of %d", what, info, ctx->M);
        SGP_FAIL(ctx, SGP_ERR_NOT_PD, buf);
    }
    return SGP_OK;
}

// In-place lower Cholesky (column-major, upper triangle zeroed); the inverses of the 64 x 64 diagonal blocks are left in
// ctx->dinv_dev.  Non-positive pivot -> SGP_ERR_NOT_PD.
int sgp_potrf_lower(sgp_ctx* ctx, double* A, int M) {
    const int nblk = (M + TB - 1) / TB;
    int rc = sgp_ensure(ctx, &ctx->dinv_dev, &ctx->dinv_cap, (size_t)nblk * TB * TB); if (rc) return rc;
    SgpDenseJob j{};
    j.M = M; j.A = A; j.Dinv = ctx->dinv_dev; j.reset_info = 1;
    rc = sgp_dense_job(ctx, j); if (rc) return rc;
    return sgp_dense_info(ctx, "potrf");
}

int sgp_gemm2(sgp_ctx* ctx, int opA, int opB, int m, int n, int k, double alpha, const double* A, int lda, const double* B, int ldb, double beta,
              double* C, int ldc, int lower_only) {
    if (m <= 0 || n <= 0) return SGP_OK;
    Gemm2Args g{A, B, C, m, n, k, lda, ldb, ldc, opA, opB, lower_only, alpha, beta};
    const size_t smem = (size_t)(STAGE_A + STAGE_B) * sizeof(double);
    const int tiles64 = ((m + 63) / 64) * ((n + 63) / 64), tiles32 = ((m + 31) / 32) * ((n + 31) / 32);
    if (tiles32 <= 2 * ctx->num_sms) gemm2_kernel<T32, 32><<<tiles32, CT, smem, ctx->stream>>>(g);
    else gemm2_kernel<T64, 64><<<std::min(tiles64, 4 * ctx->num_sms), CT, smem, ctx->stream>>>(g);
    SGP_CUDA(ctx, cudaGetLastError());
    return SGP_OK;
}

// B (M x nrhs, ld M) <- L^-1 B (trans = false) or L^-T B (trans = true) with the diagonal-block inverses `dinv` of L
// (as left by the Cholesky): per 64-wide panel one small GEMM with the inverse block and one GEMM update.
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
