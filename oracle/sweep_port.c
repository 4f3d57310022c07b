/* CPU port of the reference's data sweep, in the reference's own schedule (oracle; TEST INFRASTRUCTURE / reported CPU
 * baseline only -- never linked into libsgp.so).
 *
 * One call = what ReactiveMP does for one mini-batch of N UniSGP nodes with point-mass inputs and outputs:
 *   for each data point n (GPnode/UniSGPnode.jl:144-158):
 *       k_n   = kernelmatrix!(Psi1_trans, kernel(theta), Xu, [x_n])        M kernel evaluations
 *       Psi2  = mul!(meta.Psi2, k_n, k_n', w, 0)                            rank-1 into the shared M x M buffer
 *       xi_n  = k_n * (y_n * w)
 *     then prod (GPnode/UniSGPnode.jl:62-63):  (xi, Lambda) += (xi_n, Psi2)  M x M add
 * The flush of the N-th prod (cholinv + Cholesky) is sgp_port_flush.  Single thread, plain loops: this mirrors the
 * arithmetic the Julia code performs without the scheduler / allocation overhead, i.e. it is generous to the reference.
 * Build: gcc -O3 -fopenmp -march=native -fPIC -shared oracle/sweep_port.c -o oracle/_build/libsgp_port.so -lm
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* X: N x D (row per point), Z: M x D, Lambda: M x M (in/out), xi: M (in/out), psi2buf: M x M scratch, k: M scratch */
void sgp_port_sweep(long N, int D, int M, const double* X, const double* y, const double* Z, double variance, const double* ell,
                    double w, double* Lambda, double* xi, double* psi2buf, double* k) {
    for (long n = 0; n < N; ++n) {
        const double* x = X + n * D;
        for (int m = 0; m < M; ++m) {
            double r2 = 0.0;
            for (int d = 0; d < D; ++d) {
                double t = (x[d] - Z[(long)m * D + d]) / ell[d];
                r2 += t * t;
            }
            k[m] = variance * exp(-0.5 * r2);
        }
        for (int j = 0; j < M; ++j) {               /* mul!(Psi2, k, k', w, 0) */
            double wk = w * k[j];
            double* col = psi2buf + (long)j * M;
            for (int i = 0; i < M; ++i) col[i] = k[i] * wk;
        }
        double yw = y[n] * w;
        for (int i = 0; i < M; ++i) xi[i] += k[i] * yw;            /* prod: weighted means add */
        for (long e = 0; e < (long)M * M; ++e) Lambda[e] += psi2buf[e]; /* prod: precisions add */
    }
}

/* The same schedule with the host's threads where the reference could have them: Julia hands `mul!(Psi2, k, k', w, 0)` to a threaded
 * BLAS; here the kernel column, the rank-1 update AND the M x M add of `prod` (a single-threaded broadcast in Julia) are split over the
 * threads -- two barriers per data point.  Bitwise the same result as sgp_port_sweep (every element has one owner, same operation order).
 * Returns the number of threads used. */
#ifdef _OPENMP
#include <omp.h>
#endif
int sgp_port_sweep_mt(long N, int D, int M, const double* X, const double* y, const double* Z, double variance, const double* ell,
                      double w, double* Lambda, double* xi, double* psi2buf, double* k, int nthreads) {
    int used = 1;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#pragma omp parallel
    {
#pragma omp single
        used = omp_get_num_threads();
        for (long n = 0; n < N; ++n) {
            const double* x = X + n * D;
#pragma omp for schedule(static)
            for (int m = 0; m < M; ++m) {
                double r2 = 0.0;
                for (int d = 0; d < D; ++d) {
                    double t = (x[d] - Z[(long)m * D + d]) / ell[d];
                    r2 += t * t;
                }
                k[m] = variance * exp(-0.5 * r2);
            }   /* implicit barrier: k complete */
            const double yw = y[n] * w;
#pragma omp for schedule(static) nowait
            for (int j = 0; j < M; ++j) {
                double wk = w * k[j];
                double* col = psi2buf + (long)j * M;
                double* lam = Lambda + (long)j * M;
                for (int i = 0; i < M; ++i) col[i] = k[i] * wk;          /* mul!(Psi2, k, k', w, 0) */
                for (int i = 0; i < M; ++i) lam[i] += col[i];            /* prod: precisions add */
                xi[j] += k[j] * yw;                                      /* prod: weighted means add */
            }
#pragma omp barrier
        }
    }
#else
    (void)nthreads;
    sgp_port_sweep(N, D, M, X, y, Z, variance, ell, w, Lambda, xi, psi2buf, k);
#endif
    return used;
}

/* In-place lower Cholesky (column-major); returns 0 or the 1-based index of the failing pivot. */
static int chol_lower(double* A, int M) {
    for (int j = 0; j < M; ++j) {
        double d = A[(long)j * M + j];
        for (int l = 0; l < j; ++l) d -= A[(long)l * M + j] * A[(long)l * M + j];
        if (!(d > 0.0)) return j + 1;
        d = sqrt(d);
        A[(long)j * M + j] = d;
        for (int i = j + 1; i < M; ++i) {
            double s = A[(long)j * M + i];
            for (int l = 0; l < j; ++l) s -= A[(long)l * M + i] * A[(long)l * M + j];
            A[(long)j * M + i] = s / d;
        }
    }
    for (int j = 0; j < M; ++j)
        for (int i = 0; i < j; ++i) A[(long)j * M + i] = 0.0;
    return 0;
}

/* N-th prod (UniSGPnode.jl:65-70): Sigma = cholinv(Lambda), mu = Sigma xi, Uv = chol(Sigma + mu mu').U (returned as
 * its transpose L_R, lower, in Uv_lower).  Lambda is destroyed. */
int sgp_port_flush(int M, double* Lambda, const double* xi, double* mu, double* Sigma, double* Uv_lower) {
    int info = chol_lower(Lambda, M);
    if (info) return info;
    double* Li = (double*)calloc((size_t)M * M, sizeof(double));   /* L^-1, lower */
    for (int c = 0; c < M; ++c) {
        Li[(long)c * M + c] = 1.0 / Lambda[(long)c * M + c];
        for (int i = c + 1; i < M; ++i) {
            double s = 0.0;
            for (int l = c; l < i; ++l) s -= Lambda[(long)l * M + i] * Li[(long)c * M + l];
            Li[(long)c * M + i] = s / Lambda[(long)i * M + i];
        }
    }
    for (int j = 0; j < M; ++j)
        for (int i = j; i < M; ++i) {
            double s = 0.0;
            for (int l = i; l < M; ++l) s += Li[(long)i * M + l] * Li[(long)j * M + l];
            Sigma[(long)j * M + i] = s; Sigma[(long)i * M + j] = s;
        }
    free(Li);
    for (int i = 0; i < M; ++i) {
        double s = 0.0;
        for (int j = 0; j < M; ++j) s += Sigma[(long)j * M + i] * xi[j];
        mu[i] = s;
    }
    for (int j = 0; j < M; ++j)
        for (int i = 0; i < M; ++i) Uv_lower[(long)j * M + i] = Sigma[(long)j * M + i] + mu[i] * mu[j];
    return chol_lower(Uv_lower, M);
}
