// Uncertain inputs q(x_n) = N(m_n, S_n) (GPSSM / GPLVM): Psi statistics as expectations of the kernel column.
//
// Replaces the reference's cubature loop `approximate_kernel_expectation(!)` (GPnode/MultiSGPnode.jl:15-35,
// GPnode/UniSGPnode.jl:11-37: S closure calls per statistic per node, each allocating an M x M matrix) with
//   (a) a sigma-point generator kernel -- spherical-radial cubature, Gauss-Hermite tensor grids (ReactiveMP
//       srcubature()/ghcubature(p), restated from their published definitions) and the generalised unscented
//       transform of helper_functions/ut_approx.jl:116-151 (quirks reproduced) -- that expands the N inputs into a cloud
//       of N*S weighted virtual points, which then goes through the SAME fused generate+DMMA sweep (weighted SYRK);
//   (b) a per-point Psi1_n kernel (needed by the :out / :w rules and for vector outputs, Psi1 = sum_n Psi1_n r_n');
//   (c) closed-form SE-ARD expectations (an extension): per-point Psi2_n is not rank-1, so this
//       is an elementwise FP64-ALU-bound kernel over (pair of inducing points) x (data point).
#include "sgp_internal.cuh"
#include <cmath>
#include <vector>
#include <algorithm>

int sgp_gemm(sgp_ctx* ctx, int opA, int opB, int m, int n, int k, double alpha, const double* A, int lda, const double* B, int ldb,
             double beta, double* C, int ldc, int lower_only);

namespace {

constexpr int UD = 8;   // largest input dimension for uncertain inputs

// lower Cholesky of a d x d SPD matrix held column-major in A (in place); returns false on a non-positive pivot
__device__ bool chol_small(double* A, int d) {
    for (int j = 0; j < d; ++j) {
        double s = A[j + j * d];
        for (int l = 0; l < j; ++l) s -= A[j + l * d] * A[j + l * d];
        if (!(s > 0.0)) return false;
        s = sqrt(s);
        A[j + j * d] = s;
        for (int i = j + 1; i < d; ++i) {
            double t = A[i + j * d];
            for (int l = 0; l < j; ++l) t -= A[i + l * d] * A[j + l * d];
            A[i + j * d] = t / s;
        }
        for (int i = 0; i < j; ++i) A[i + j * d] = 0.0;
    }
    return true;
}

// One thread per input point: writes S sigma points (point-major, d doubles each), S weights and S copies of r_n.
__global__ void sigma_kernel(int method, int p, int d, long long N, int S, const double* __restrict__ mean, const double* __restrict__ cov,
                             const double* __restrict__ r, const double* __restrict__ gh, double* __restrict__ Xv, double* __restrict__ wv,
                             double* __restrict__ yv, int* __restrict__ info) {
    long long n = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (n >= N) return;
    double L[UD * UD], m[UD];
    for (int i = 0; i < d; ++i) m[i] = mean[n * d + i];
    for (int i = 0; i < d * d; ++i) L[i] = (method == SGP_METHOD_POINT) ? 0.0 : cov[n * d * d + i];
    // symmetrise from both triangles (the reference hands a Hermitian view to the factorisation)
    for (int i = 0; i < d; ++i)
        for (int j = 0; j < i; ++j) { double v = 0.5 * (L[i + j * d] + L[j + i * d]); L[i + j * d] = v; L[j + i * d] = v; }
    double V00 = L[0];
    double* X = Xv + n * S * d;
    double* w = wv + n * S;
    const double rn = r ? r[n] : 1.0;
    for (int s = 0; s < S; ++s) yv[n * S + s] = rn;
    if (method == SGP_METHOD_POINT) {       // q(x_n) = delta(m_n): one point, weight 1 (the covariance is not read)
        for (int i = 0; i < d; ++i) X[i] = m[i];
        w[0] = 1.0;
        return;
    }
    if (!chol_small(L, d)) { atomicExch(info, 1); return; }
    if (method == SGP_METHOD_SRCUBATURE) {
        // m +/- sqrt(d+1) L e_j, weight 1/(2(d+1)); centre last, weight 1/(d+1)
        const double sc = sqrt((double)d + 1.0);
        for (int j = 0; j < d; ++j)
            for (int i = 0; i < d; ++i) {
                X[j * d + i] = m[i] + sc * L[i + j * d];
                X[(d + j) * d + i] = m[i] - sc * L[i + j * d];
            }
        for (int i = 0; i < d; ++i) X[2 * d * d + i] = m[i];
        for (int s = 0; s < 2 * d; ++s) w[s] = 1.0 / (2.0 * (d + 1.0));
        w[2 * d] = 1.0 / (d + 1.0);
    } else if (method == SGP_METHOD_GENUT) {
        if (d == 1) {
            // helper_functions/ut_approx.jl:116-126 with S = 0, K = 3 (Normal): u = v = sqrt(12)/(2V)
            const double V = V00, Ls = sqrt(V);
            const double u = 0.5 * (1.0 / V) * sqrt(12.0), v = u;
            const double aux = 1.0 / (v * (u + v));
            X[0] = m[0]; X[1] = m[0] - u * Ls; X[2] = m[0] + v * Ls;
            w[0] = 1.0 - aux * (v / u + 1.0); w[1] = (v / u) * aux; w[2] = aux;
        } else {
            // helper_functions/ut_approx.jl:129-151: only diag(L)^3 survives cholinv(L.^3) (upper-triangle factorisation)
            for (int i = 0; i < d; ++i) X[i] = m[i];
            double sumw = 0.0;
            for (int j = 0; j < d; ++j) {
                double l3 = L[j + j * d] * L[j + j * d] * L[j + j * d];
                double u = 0.5 * sqrt(12.0 / (l3 * l3)), v = u;
                for (int i = 0; i < d; ++i) {
                    X[(1 + j) * d + i] = m[i] - L[i + j * d] * u;
                    X[(1 + d + j) * d + i] = m[i] + L[i + j * d] * v;
                }
                double wp = 1.0 / v / (u + v);
                w[1 + d + j] = wp; w[1 + j] = wp * (v / u);
                sumw += w[1 + d + j] + w[1 + j];
            }
            w[0] = 1.0 - sumw;
        }
    } else {   // Gauss-Hermite tensor grid, first index slowest (matches oracle/cubature.py meshgrid 'ij')
        const double* t = gh; const double* gw = gh + p;
        const double s2 = 1.4142135623730951;
        double norm = 1.0;
        for (int i = 0; i < d; ++i) norm *= 0.5641895835477563;    // 1/sqrt(pi) per dimension
        for (int s = 0; s < S; ++s) {
            int idx[UD]; int q = s;
            for (int i = d - 1; i >= 0; --i) { idx[i] = q % p; q /= p; }
            double ww = norm;
            for (int i = 0; i < d; ++i) ww *= gw[idx[i]];
            for (int i = 0; i < d; ++i) {
                double a = m[i];
                for (int j = 0; j <= i; ++j) a = fma(s2 * L[i + j * d], t[idx[j]], a);
                X[s * d + i] = a;
            }
            w[s] = ww;
        }
    }
}

__device__ __forceinline__ double kernel_val(int kind, double r2) {
    if (kind == SGP_KERNEL_SE) return exp(-0.5 * r2);
    double s = sqrt((kind == SGP_KERNEL_MATERN32 ? 3.0 : 5.0) * r2);
    return kind == SGP_KERNEL_MATERN32 ? (1.0 + s) * exp(-s) : (1.0 + s + s * s / 3.0) * exp(-s);
}

// psi1_n[m + n*M] = sum_s w_s k(x_ns, z_m): one thread per (m, n)
__global__ void psi1n_cloud_kernel(const double* __restrict__ Xv, const double* __restrict__ wv, const double* __restrict__ Z,
                                   const double* __restrict__ ell_inv, double* __restrict__ out, long long N, int S, int M, int d, int kind,
                                   double variance) {
    long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (e >= N * M) return;
    long long n = e / M; int m = (int)(e - n * M);
    double acc = 0.0;
    for (int s = 0; s < S; ++s) {
        double r2 = 0.0;
        for (int i = 0; i < d; ++i) { double t = (Xv[(n * S + s) * d + i] - Z[(size_t)m * d + i]) * ell_inv[i]; r2 = fma(t, t, r2); }
        acc = fma(wv[n * S + s], kernel_val(kind, r2), acc);
    }
    out[e] = variance * acc;
}

// ---- closed-form SE-ARD ---------------------------------------------------------------------------------------------
// per-point record: [B1 (d*d) | c1 | B2 (d*d) | c2 | m (d)]
__device__ bool inv_spd_small(const double* A, int d, double* Ainv, double* logdet) {
    double L[UD * UD];
    for (int i = 0; i < d * d; ++i) L[i] = A[i];
    if (!chol_small(L, d)) return false;
    double ld = 0.0;
    for (int i = 0; i < d; ++i) ld += 2.0 * log(L[i + i * d]);
    *logdet = ld;
    double Li[UD * UD];
    for (int i = 0; i < d * d; ++i) Li[i] = 0.0;
    for (int c = 0; c < d; ++c) {
        Li[c + c * d] = 1.0 / L[c + c * d];
        for (int i = c + 1; i < d; ++i) {
            double s = 0.0;
            for (int l = c; l < i; ++l) s -= L[i + l * d] * Li[l + c * d];
            Li[i + c * d] = s / L[i + i * d];
        }
    }
    for (int i = 0; i < d; ++i)
        for (int j = 0; j < d; ++j) {
            double s = 0.0;
            for (int l = (i > j ? i : j); l < d; ++l) s += Li[l + i * d] * Li[l + j * d];
            Ainv[i + j * d] = s;
        }
    return true;
}

__global__ void cf_prep_kernel(int d, long long N, const double* __restrict__ mean, const double* __restrict__ cov, const double* __restrict__ ell,
                               double variance, double* __restrict__ rec, int* __restrict__ info) {
    long long n = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (n >= N) return;
    const int RS = 2 * d * d + 2 + d;
    double* r = rec + n * RS;
    double A[UD * UD], Sm[UD * UD];
    double logdetLam = 0.0;
    for (int i = 0; i < d; ++i) logdetLam += 2.0 * log(ell[i]);
    for (int i = 0; i < d; ++i)
        for (int j = 0; j < d; ++j) Sm[i + j * d] = 0.5 * (cov[n * d * d + i + j * d] + cov[n * d * d + j + i * d]);
    double ld;
    for (int pass = 0; pass < 2; ++pass) {
        const double f = pass == 0 ? 1.0 : 2.0;
        for (int i = 0; i < d * d; ++i) A[i] = f * Sm[i];
        for (int i = 0; i < d; ++i) A[i + i * d] += ell[i] * ell[i];
        double* B = r + pass * (d * d + 1);
        if (!inv_spd_small(A, d, B, &ld)) { atomicExch(info, 1); return; }
        // |I + f Lam^-1 S|^(-1/2) = exp(-(logdet(Lam + f S) - logdet Lam)/2)
        B[d * d] = (pass == 0 ? variance : variance * variance) * exp(-0.5 * (ld - logdetLam));
    }
    for (int i = 0; i < d; ++i) r[2 * d * d + 2 + i] = mean[n * d + i];
}

__global__ void cf_psi1n_kernel(const double* __restrict__ rec, const double* __restrict__ Z, double* __restrict__ out, long long N, int M, int d) {
    long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (e >= N * M) return;
    long long n = e / M; int a = (int)(e - n * M);
    const int RS = 2 * d * d + 2 + d;
    const double* r = rec + n * RS;
    double dm[UD];
    for (int i = 0; i < d; ++i) dm[i] = r[2 * d * d + 2 + i] - Z[(size_t)a * d + i];
    double q = 0.0;
    for (int i = 0; i < d; ++i) {
        double s = 0.0;
        for (int j = 0; j < d; ++j) s = fma(r[i + j * d], dm[j], s);
        q = fma(dm[i], s, q);
    }
    out[e] = r[d * d] * exp(-0.5 * q);
}

// partial[split][a + b*M] (a >= b) = sum over the split's points of c2_n exp(-(m_n - zbar)' B2_n (m_n - zbar))
__global__ void __launch_bounds__(256) cf_psi2_kernel(const double* __restrict__ rec, const double* __restrict__ Z, double* __restrict__ partial,
                                                      long long N, int M, int d, int nsplit) {
    extern __shared__ double sh[];
    const int RS = 2 * d * d + 2 + d;
    const int PC = 64;                       // points staged per round
    long long pair = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long npairs = (long long)M * (M + 1) / 2;
    int a = 0, b = 0;
    const bool active = pair < npairs;
    if (active) {
        a = (int)((sqrt(8.0 * (double)pair + 1.0) - 1.0) * 0.5);
        while ((long long)(a + 1) * (a + 2) / 2 <= pair) ++a;
        while ((long long)a * (a + 1) / 2 > pair) --a;
        b = (int)(pair - (long long)a * (a + 1) / 2);
    }
    double zb[UD];
    for (int i = 0; i < d; ++i) zb[i] = active ? 0.5 * (Z[(size_t)a * d + i] + Z[(size_t)b * d + i]) : 0.0;
    const int split = blockIdx.y;
    const long long n0 = N * split / nsplit, n1 = N * (split + 1) / nsplit;
    double acc = 0.0;
    for (long long base = n0; base < n1; base += PC) {
        int cnt = (int)((n1 - base) < PC ? (n1 - base) : PC);
        __syncthreads();
        for (int e = threadIdx.x; e < cnt * RS; e += blockDim.x) sh[e] = rec[base * RS + e];
        __syncthreads();
        if (active)
            for (int q = 0; q < cnt; ++q) {
                const double* r = sh + q * RS;
                const double* B = r + d * d + 1;
                double dm[UD];
                for (int i = 0; i < d; ++i) dm[i] = r[2 * d * d + 2 + i] - zb[i];
                double qf = 0.0;
                for (int i = 0; i < d; ++i) {
                    double s = 0.0;
                    for (int j = 0; j < d; ++j) s = fma(B[i + j * d], dm[j], s);
                    qf = fma(dm[i], s, qf);
                }
                acc = fma(B[d * d], exp(-qf), acc);
            }
    }
    if (active) partial[(size_t)split * M * M + a + (size_t)b * M] = acc;
}

// Closed-form Psi2, fast path (d <= 4): per-point packed record [s*B2 upper-packed with doubled off-diagonals | c2 | m],
// s = -2048/ln2, so that the quadratic form IS the argument of the table-driven exp (sgp_internal.cuh: 7 FP64 instructions);
// one thread per inducing pair (a >= b), points staged through shared memory, four points in flight per thread.
// FP64-ALU-bound: d + d(d+1)/2 + d ... ~ 15 FP64 instructions per (pair, point) at d = 2.
template <int DD>
__global__ void __launch_bounds__(256) cf_psi2_fast_kernel(const double* __restrict__ rec2, const double* __restrict__ Z, const double* __restrict__ exptab,
                                                           double* __restrict__ partial, long long N, int M, int nsplit) {
    constexpr int NP = DD * (DD + 1) / 2, RS = NP + 1 + DD, PC = 128;
    __shared__ double tab[SGP_EXP_TAB];
    __shared__ double sh[PC * RS];
    for (int i = threadIdx.x; i < SGP_EXP_TAB; i += 256) tab[i] = exptab[i];
    const long long pair = blockIdx.x * 256LL + threadIdx.x;
    const long long npairs = (long long)M * (M + 1) / 2;
    int a = 0, b = 0;
    const bool active = pair < npairs;
    if (active) {
        a = (int)((sqrt(8.0 * (double)pair + 1.0) - 1.0) * 0.5);
        while ((long long)(a + 1) * (a + 2) / 2 <= pair) ++a;
        while ((long long)a * (a + 1) / 2 > pair) --a;
        b = (int)(pair - (long long)a * (a + 1) / 2);
    }
    double zb[DD];
#pragma unroll
    for (int i = 0; i < DD; ++i) zb[i] = active ? 0.5 * (Z[(size_t)a * DD + i] + Z[(size_t)b * DD + i]) : 0.0;
    const int split = blockIdx.y;
    const long long n0 = N * split / nsplit, n1 = N * (split + 1) / nsplit;
    double acc = 0.0;
    for (long long base = n0; base < n1; base += PC) {
        const int cnt = (int)((n1 - base) < PC ? (n1 - base) : PC);
        __syncthreads();
        for (int e = threadIdx.x; e < PC * RS; e += 256) sh[e] = (e < cnt * RS) ? rec2[base * RS + e] : 0.0;   // padded points: c2 = 0
        __syncthreads();
#pragma unroll 1
        for (int q0 = 0; q0 < PC; q0 += 4) {
            if (q0 >= cnt) break;
            double t[4], c2[4], k[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const double* r = sh + (q0 + u) * RS;
                double dm[DD];
#pragma unroll
                for (int i = 0; i < DD; ++i) dm[i] = r[NP + 1 + i] - zb[i];
                double qf = 0.0;
                int p = 0;
#pragma unroll
                for (int i = 0; i < DD; ++i) {
                    double s = 0.0;
#pragma unroll
                    for (int j = i; j < DD; ++j, ++p) s = fma(r[p], dm[j], s);
                    qf = fma(dm[i], s, qf);
                }
                t[u] = qf; c2[u] = r[NP];
            }
            exp_scaled_v<4>(t, k, tab);
#pragma unroll
            for (int u = 0; u < 4; ++u) acc = fma(c2[u], k[u], acc);
        }
    }
    if (active) partial[(size_t)split * M * M + a + (size_t)b * M] = acc;
}

// rec (generic record of cf_prep_kernel) -> packed, pre-scaled Psi2 record
__global__ void cf_pack_kernel(const double* __restrict__ rec, double* __restrict__ rec2, long long N, int d) {
    const long long n = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (n >= N) return;
    const int RS = 2 * d * d + 2 + d, NP = d * (d + 1) / 2, RS2 = NP + 1 + d;
    const double* r = rec + n * RS;
    const double* B = r + d * d + 1;
    double* o = rec2 + n * RS2;
    int p = 0;
    for (int i = 0; i < d; ++i)
        for (int j = i; j < d; ++j, ++p) o[p] = -SGP_EXP_SCALE * (i == j ? B[i + j * d] : B[i + j * d] + B[j + i * d]);
    o[NP] = B[d * d];
    for (int i = 0; i < d; ++i) o[NP + 1 + i] = r[2 * d * d + 2 + i];
}

__global__ void cf_finish_kernel(const double* __restrict__ partial, const double* __restrict__ Z, const double* __restrict__ ell_inv,
                                 double* __restrict__ psi2, int M, int d, int nsplit) {
    size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (e >= (size_t)M * M) return;
    int a = (int)(e % M), b = (int)(e / M);
    if (a < b) return;
    double v = 0.0;
    for (int s = 0; s < nsplit; ++s) v += partial[(size_t)s * M * M + e];
    double q = 0.0;
    for (int i = 0; i < d; ++i) { double t = (Z[(size_t)a * d + i] - Z[(size_t)b * d + i]) * ell_inv[i]; q = fma(t, t, q); }
    v *= exp(-0.25 * q);
    psi2[(size_t)a + (size_t)b * M] = v;
    psi2[(size_t)b + (size_t)a * M] = v;
}

// Psi1 (M x D_out) = Psi1_n (M x N) R (N x D_out): the output is tiny and K = N is huge -> split N over blocks, one thread per
// inducing row (coalesced), fixed-order finish (deterministic)
template <int DO>
__global__ void __launch_bounds__(128) psi1_split_kernel(const double* __restrict__ p1n, const double* __restrict__ R, double* __restrict__ partial,
                                                         long long N, int M, int nsplit) {
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    const int split = blockIdx.y;
    const long long n0 = N * split / nsplit, n1 = N * (split + 1) / nsplit;
    double acc[DO];
#pragma unroll
    for (int d = 0; d < DO; ++d) acc[d] = 0.0;
    if (m < M)
        for (long long n = n0; n < n1; ++n) {
            const double v = p1n[(size_t)m + (size_t)n * M];
#pragma unroll
            for (int d = 0; d < DO; ++d) acc[d] = fma(v, R[(size_t)n + (size_t)d * N], acc[d]);
        }
    if (m < M)
#pragma unroll
        for (int d = 0; d < DO; ++d) partial[((size_t)split * DO + d) * M + m] = acc[d];
}
__global__ void psi1_finish_kernel(const double* __restrict__ partial, double* __restrict__ psi1, int M, int DO, int nsplit) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= M * DO) return;
    const int m = e % M, d = e / M;
    double v = 0.0;
    for (int s = 0; s < nsplit; ++s) v += partial[((size_t)s * DO + d) * M + m];
    psi1[(size_t)m + (size_t)d * M] = v;
}

__global__ void set_scal_kernel(double* scal, double psi0, double n) { scal[0] = psi0; scal[1] = 0.0; scal[2] = n; scal[3] = n; }
__global__ void copy4_kernel(double* dst, const double* src) { if (threadIdx.x < 4) dst[threadIdx.x] = src[threadIdx.x]; }

void gauss_hermite(int p, std::vector<double>& t, std::vector<double>& w) {
    // Newton on the orthonormal Hermite recurrence (physicists' weight exp(-x^2)); accurate to ~1e-15 for p <= 64
    t.assign(p, 0.0); w.assign(p, 0.0);
    const double pim4 = 0.7511255444649425;
    int mhalf = (p + 1) / 2;
    double z = 0.0, pp = 0.0;
    for (int i = 0; i < mhalf; ++i) {
        if (i == 0) z = std::sqrt(2.0 * p + 1.0) - 1.85575 * std::pow(2.0 * p + 1.0, -0.16667);
        else if (i == 1) z -= 1.14 * std::pow((double)p, 0.426) / z;
        else if (i == 2) z = 1.86 * z - 0.86 * t[0];
        else if (i == 3) z = 1.91 * z - 0.91 * t[1];
        else z = 2.0 * z - t[i - 2];
        for (int it = 0; it < 100; ++it) {
            double p1 = pim4, p2 = 0.0;
            for (int j = 0; j < p; ++j) { double p3 = p2; p2 = p1; p1 = z * std::sqrt(2.0 / (j + 1)) * p2 - std::sqrt((double)j / (j + 1)) * p3; }
            pp = std::sqrt(2.0 * p) * p2;
            double z1 = z; z = z1 - p1 / pp;
            if (std::fabs(z - z1) <= 1e-16 * (1.0 + std::fabs(z))) break;
        }
        t[i] = z; t[p - 1 - i] = -z;
        w[i] = 2.0 / (pp * pp); w[p - 1 - i] = w[i];
    }
    // ascending order like numpy.polynomial.hermite.hermgauss
    for (int i = 0; i < p / 2; ++i) { std::swap(t[i], t[p - 1 - i]); std::swap(w[i], w[p - 1 - i]); }
}

inline unsigned nb(size_t n, int t = 256) { return (unsigned)((n + t - 1) / t); }

}  // namespace

int sgp_uncertain_sweep(sgp_ctx* ctx, int method, int p, int64_t N, const double* mean, const double* cov, int D_out, const double* R,
                        double* psi0, double* psi1, double* psi2, double* psi1_n) {
    if (!ctx->have_kernel || !ctx->have_Z) SGP_FAIL(ctx, SGP_ERR_ARG, "sweep_uncertain: set_kernel and set_inducing first");
    if (N < 1 || !mean || (!cov && method != SGP_METHOD_POINT) || D_out < 1) SGP_FAIL(ctx, SGP_ERR_ARG, "sweep_uncertain: bad arguments");
    const int d = ctx->D, M = ctx->M;
    if (d > UD) SGP_FAIL(ctx, SGP_ERR_UNSUPPORTED, "sweep_uncertain: input dimension <= 8");
    if (method == SGP_METHOD_CLOSED_FORM_SE && ctx->kind != SGP_KERNEL_SE) SGP_FAIL(ctx, SGP_ERR_UNSUPPORTED, "closed form exists for SE-ARD only");
    int S = 0;
    if (method == SGP_METHOD_SRCUBATURE) S = 2 * d + 1;
    else if (method == SGP_METHOD_GENUT) S = d == 1 ? 3 : 2 * d + 1;
    else if (method == SGP_METHOD_GAUSSHERMITE) {
        if (p < 1 || p > 64) SGP_FAIL(ctx, SGP_ERR_ARG, "Gauss-Hermite order 1..64");
        double s = 1.0; for (int i = 0; i < d; ++i) s *= p;
        if (s > 4096.0) SGP_FAIL(ctx, SGP_ERR_UNSUPPORTED, "Gauss-Hermite tensor grid larger than 4096 points per input");
        S = (int)s;
    } else if (method == SGP_METHOD_POINT) S = 1;
    else if (method != SGP_METHOD_CLOSED_FORM_SE) SGP_FAIL(ctx, SGP_ERR_ARG, "unknown method");

    const bool need_p1n = psi1_n != nullptr || D_out > 1 || method == SGP_METHOD_CLOSED_FORM_SE;
    const size_t MM = (size_t)M * M;
    // device scratch: inputs
    double *mean_d = nullptr, *cov_d = nullptr, *R_d = nullptr, *p1n_d = nullptr, *misc_d = nullptr;
    int rc = SGP_OK;
    // all scratch of the call is carved from ONE context-owned arena that only ever grows: no cudaMalloc / cudaFree (and the
    // implicit device synchronisations they bring) per call
    const int RS_cf = 2 * d * d + 2 + d, RS2_cf = d * (d + 1) / 2 + 1 + d;
    const int psi1_split = (int)std::max<long long>(1, std::min<long long>(512, N / 64));
    size_t arena_need = (size_t)N * d + (size_t)N * d * d + 512 + (size_t)N * (size_t)std::max(D_out, 1) + (size_t)N
                      + (need_p1n ? (size_t)N * M : 0) + (method == SGP_METHOD_CLOSED_FORM_SE ? (size_t)N * (RS_cf + RS2_cf) : 0)
                      + (size_t)psi1_split * std::max(D_out, 1) * M + 64;
    rc = sgp_ensure(ctx, &ctx->unc_dev, &ctx->unc_cap, arena_need); if (rc) return rc;
    size_t arena_off = 0;
    auto take = [&](size_t n) { double* p = ctx->unc_dev + arena_off; arena_off += (n + 1) & ~(size_t)1; return p; };
    auto cleanup = [&]() {};
#define UC(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { ctx->err = std::string(#call) + ": " + cudaGetErrorString(e__); cleanup(); return SGP_ERR_CUDA; } } while (0)
    mean_d = take((size_t)N * d);
    cov_d = take((size_t)N * d * d);
    misc_d = take(256 + SGP_MAX_D);
    UC(cudaMemcpyAsync(mean_d, mean, (size_t)N * d * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    if (cov) UC(cudaMemcpyAsync(cov_d, cov, (size_t)N * d * d * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    if (R) {
        R_d = take((size_t)N * D_out);
        UC(cudaMemcpyAsync(R_d, R, (size_t)N * D_out * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    }
    if (need_p1n) p1n_d = take((size_t)N * M);
    double hostmisc[256 + SGP_MAX_D] = {0};
    for (int i = 0; i < d; ++i) hostmisc[i] = 1.0 / ctx->ell[i];
    for (int i = 0; i < d; ++i) hostmisc[SGP_MAX_D + i] = ctx->ell[i];
    if (method == SGP_METHOD_GAUSSHERMITE) {
        std::vector<double> t, w; gauss_hermite(p, t, w);
        for (int i = 0; i < p; ++i) { hostmisc[2 * SGP_MAX_D + i] = t[i]; hostmisc[2 * SGP_MAX_D + p + i] = w[i]; }
    }
    UC(cudaMemcpyAsync(misc_d, hostmisc, sizeof hostmisc, cudaMemcpyHostToDevice, ctx->stream));
    UC(cudaMemsetAsync(ctx->info_dev, 0, sizeof(int), ctx->stream));
    const double* ell_inv_d = misc_d; const double* ell_d = misc_d + SGP_MAX_D; const double* gh_d = misc_d + 2 * SGP_MAX_D;

    rc = sgp_ensure_stats(ctx, MM + (size_t)M * (D_out > 1 ? D_out : 1) + 8);
    if (rc) { cleanup(); return rc; }
    double* s_psi2 = ctx->stats_dev; double* s_psi1 = s_psi2 + MM; double* s_scal = s_psi1 + (size_t)M * D_out;
    ctx->packed_src = nullptr;       // the resident statistics are rewritten below: a packed copy of the previous ones is stale

    if (method != SGP_METHOD_CLOSED_FORM_SE) {
        const size_t NS = (size_t)N * S, cap = ((NS + 31) / 32) * 32;
        if (ctx->sp_cap < cap) {
            cudaFree(ctx->sp_X_dev); cudaFree(ctx->sp_w_dev); cudaFree(ctx->sp_y_dev); ctx->sp_X_dev = ctx->sp_w_dev = ctx->sp_y_dev = nullptr; ctx->sp_cap = 0;
            UC(cudaMalloc((void**)&ctx->sp_X_dev, cap * UD * sizeof(double)));   // sized for the largest input dimension: the buffer is reused across calls
            UC(cudaMalloc((void**)&ctx->sp_w_dev, cap * sizeof(double)));
            UC(cudaMalloc((void**)&ctx->sp_y_dev, cap * sizeof(double)));
            ctx->sp_cap = cap;
        }
        UC(cudaMemsetAsync(ctx->sp_X_dev, 0, cap * d * sizeof(double), ctx->stream));
        UC(cudaMemsetAsync(ctx->sp_w_dev, 0, cap * sizeof(double), ctx->stream));
        UC(cudaMemsetAsync(ctx->sp_y_dev, 0, cap * sizeof(double), ctx->stream));
        sigma_kernel<<<nb((size_t)N, 128), 128, 0, ctx->stream>>>(method, p, d, N, S, mean_d, cov_d, (R_d && D_out == 1) ? R_d : nullptr, gh_d,
                                                                 ctx->sp_X_dev, ctx->sp_w_dev, ctx->sp_y_dev, ctx->info_dev);
        UC(cudaGetLastError());
        ctx->sp_N = N; ctx->sp_S = S;          // the cloud stays resident for sgp_uncertain_node_terms
        // weighted fused sweep over the cloud: Psi2 = sum w k k', Psi1 (D_out = 1) = sum w r k, Psi0 = sigma^2 sum w
        rc = sgp_sweep_launch(ctx, ctx->sp_X_dev, ctx->sp_y_dev, nullptr, ctx->sp_w_dev, (int64_t)NS, (int64_t)cap, false);
        if (rc) { cleanup(); return rc; }
        if (D_out > 1) copy4_kernel<<<1, 32, 0, ctx->stream>>>(s_scal, s_psi2 + MM + M);      // scalars move behind the wider Psi1
        if (need_p1n)
            psi1n_cloud_kernel<<<nb((size_t)N * M), 256, 0, ctx->stream>>>(ctx->sp_X_dev, ctx->sp_w_dev, ctx->Z_dev, ell_inv_d, p1n_d, N, S, M, d,
                                                                           ctx->kind, ctx->variance);
    } else {
        ctx->sp_N = 0; ctx->sp_S = 0;          // closed form: no sigma-point cloud
        const int RS = 2 * d * d + 2 + d;
        double* rec_d = nullptr;
        rec_d = take((size_t)N * RS);
        cf_prep_kernel<<<nb((size_t)N, 128), 128, 0, ctx->stream>>>(d, N, mean_d, cov_d, ell_d, ctx->variance, rec_d, ctx->info_dev);
        cf_psi1n_kernel<<<nb((size_t)N * M), 256, 0, ctx->stream>>>(rec_d, ctx->Z_dev, p1n_d, N, M, d);
        const long long npairs = (long long)M * (M + 1) / 2;
        int nsplit = (int)std::max<long long>(1, std::min<long long>(64, (4LL * ctx->num_sms * 256) / std::max<long long>(npairs, 1)));
        if (nsplit > N) nsplit = (int)N;
        rc = sgp_ensure(ctx, &ctx->work_dev, &ctx->work_cap, (size_t)nsplit * MM);
        if (rc) { cleanup(); return rc; }
        cudaMemsetAsync(ctx->work_dev, 0, (size_t)nsplit * MM * sizeof(double), ctx->stream);
        dim3 grid(nb((size_t)npairs), nsplit);
        if (d <= 4) {
            double* rec2_d = nullptr;
            const int RS2 = d * (d + 1) / 2 + 1 + d;
            rec2_d = take((size_t)N * RS2);
            cf_pack_kernel<<<nb((size_t)N, 128), 128, 0, ctx->stream>>>(rec_d, rec2_d, N, d);
            switch (d) {
                case 1: cf_psi2_fast_kernel<1><<<grid, 256, 0, ctx->stream>>>(rec2_d, ctx->Z_dev, ctx->exptab_dev, ctx->work_dev, N, M, nsplit); break;
                case 2: cf_psi2_fast_kernel<2><<<grid, 256, 0, ctx->stream>>>(rec2_d, ctx->Z_dev, ctx->exptab_dev, ctx->work_dev, N, M, nsplit); break;
                case 3: cf_psi2_fast_kernel<3><<<grid, 256, 0, ctx->stream>>>(rec2_d, ctx->Z_dev, ctx->exptab_dev, ctx->work_dev, N, M, nsplit); break;
                default: cf_psi2_fast_kernel<4><<<grid, 256, 0, ctx->stream>>>(rec2_d, ctx->Z_dev, ctx->exptab_dev, ctx->work_dev, N, M, nsplit); break;
            }
        } else {
            cf_psi2_kernel<<<grid, 256, 64 * RS * sizeof(double), ctx->stream>>>(rec_d, ctx->Z_dev, ctx->work_dev, N, M, d, nsplit);
        }
        cf_finish_kernel<<<nb(MM), 256, 0, ctx->stream>>>(ctx->work_dev, ctx->Z_dev, ell_inv_d, s_psi2, M, d, nsplit);
        set_scal_kernel<<<1, 1, 0, ctx->stream>>>(s_scal, ctx->variance * (double)N, (double)N);
        cudaError_t e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) { ctx->err = std::string("closed-form kernels: ") + cudaGetErrorString(e); cleanup(); return SGP_ERR_CUDA; }
        ctx->last_launches = 5;
    }
    ctx->Dout = D_out; ctx->have_stats = true;
    if (need_p1n) {
        // Psi1 (M x D_out) = Psi1_n (M x N) * R (N x D_out); R == NULL -> ones
        if (!R_d) {
            R_d = take((size_t)N);
            std::vector<double> ones((size_t)N, 1.0);
            UC(cudaMemcpyAsync(R_d, ones.data(), (size_t)N * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
            UC(cudaStreamSynchronize(ctx->stream));
        }
        if (D_out <= 4) {
            const int nsplit = psi1_split;
            double* part = take((size_t)nsplit * D_out * M);
            dim3 g((M + 127) / 128, nsplit);
            switch (D_out) {
                case 1: psi1_split_kernel<1><<<g, 128, 0, ctx->stream>>>(p1n_d, R_d, part, N, M, nsplit); break;
                case 2: psi1_split_kernel<2><<<g, 128, 0, ctx->stream>>>(p1n_d, R_d, part, N, M, nsplit); break;
                case 3: psi1_split_kernel<3><<<g, 128, 0, ctx->stream>>>(p1n_d, R_d, part, N, M, nsplit); break;
                default: psi1_split_kernel<4><<<g, 128, 0, ctx->stream>>>(p1n_d, R_d, part, N, M, nsplit); break;
            }
            psi1_finish_kernel<<<nb((size_t)M * D_out), 256, 0, ctx->stream>>>(part, s_psi1, M, D_out, nsplit);
        } else {
            rc = sgp_gemm(ctx, 0, 0, M, D_out, (int)N, 1.0, p1n_d, M, R_d, (int)N, 0.0, s_psi1, M, 0);
            if (rc) { cleanup(); return rc; }
        }
    }
    // N sharded over ranks: the statistics are sums over the nodes of ALL ranks (the per-node Psi1_n stay local) -- COLLECTIVE like sgp_sweep_psi
    if (ctx->comm) { rc = sgp_comm_allreduce_stats(ctx, M, D_out); if (rc) { cleanup(); return rc; } }
    int info = 0;
    UC(cudaMemcpyAsync(&info, ctx->info_dev, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    double sc[4];
    UC(cudaMemcpyAsync(sc, s_scal, sizeof sc, cudaMemcpyDeviceToHost, ctx->stream));
    if (psi2) UC(cudaMemcpyAsync(psi2, s_psi2, MM * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    if (psi1) UC(cudaMemcpyAsync(psi1, s_psi1, (size_t)M * D_out * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    if (psi1_n) UC(cudaMemcpyAsync(psi1_n, p1n_d, (size_t)N * M * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    UC(cudaStreamSynchronize(ctx->stream));
    cleanup();
#undef UC
    if (info) SGP_FAIL(ctx, SGP_ERR_NOT_PD, "sweep_uncertain: an input covariance is not positive definite");
    if (psi0) *psi0 = sc[0];
    return SGP_OK;
}
