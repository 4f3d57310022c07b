/* A plain-C client of libsgp.so: the drop-in boundary used without Python, torch or C++ -- what a `ccall` from Julia amounts to.
 * One regression sweep + the N-th `prod` (GPnode/UniSGPnode.jl:144-158, 62-73) on a small deterministic problem; prints the results
 * as text so that a test can compare them with the oracle.
 *   build:  gcc -std=c11 -Wall -Werror -Iinclude examples/c_client.c -o c_client -Lgaussianprocessnode_b200 -lsgp -lm -Wl,-rpath,$PWD/gaussianprocessnode_b200
 *   run:    ./c_client [N M D]            (needs a CUDA device: there is no CPU fallback) */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include "sgp.h"

static double lcg(unsigned long long* s) {            /* deterministic inputs without any library RNG: uniform in (-1, 1) */
    *s = *s * 6364136223846793005ULL + 1442695040888963407ULL;
    return (double)((*s >> 11) & ((1ULL << 53) - 1)) / (double)(1ULL << 52) - 1.0;
}

#define CK(call) do { int rc_ = (call); if (rc_ != SGP_OK) { fprintf(stderr, "%s -> %d: %s\n", #call, rc_, ctx ? sgp_last_error(ctx) : "no context"); return 1; } } while (0)

int main(int argc, char** argv) {
    const int N = argc > 3 ? atoi(argv[1]) : 200, M = argc > 3 ? atoi(argv[2]) : 24, D = argc > 3 ? atoi(argv[3]) : 3;
    sgp_ctx* ctx = NULL;
    unsigned long long seed = 12345;
    double *X = malloc(sizeof(double) * N * D), *y = malloc(sizeof(double) * N), *Z = malloc(sizeof(double) * M * D);
    double *ell = malloc(sizeof(double) * D), *psi1 = malloc(sizeof(double) * M), *psi2 = malloc(sizeof(double) * M * M);
    double *xi0 = calloc(M, sizeof(double)), *Lam0 = calloc((size_t)M * M, sizeof(double));
    double *mu = malloc(sizeof(double) * M), *Sig = malloc(sizeof(double) * M * M), *Uv = malloc(sizeof(double) * M * M);
    double psi0 = 0.0, sy2 = 0.0, s1 = 0.0, s2 = 0.0;
    if (!X || !y || !Z || !ell || !psi1 || !psi2 || !xi0 || !Lam0 || !mu || !Sig || !Uv) return 2;
    for (int i = 0; i < N * D; ++i) X[i] = 2.0 * lcg(&seed);
    for (int n = 0; n < N; ++n) y[n] = sin(X[(size_t)n * D]) + 0.05 * lcg(&seed);
    for (int i = 0; i < M * D; ++i) Z[i] = 2.0 * lcg(&seed);
    for (int d = 0; d < D; ++d) ell[d] = 1.0 + 0.1 * d;
    for (int m = 0; m < M; ++m) Lam0[(size_t)m * M + m] = 1.0 / 50.0;      /* prior N(0, 50 I), regression_kin40k.ipynb:200-201 */

    CK(sgp_create(&ctx, 0));
    CK(sgp_set_kernel(ctx, SGP_KERNEL_SE, D, 1.3, ell));
    CK(sgp_set_inducing(ctx, M, Z));
    CK(sgp_sweep_psi_host(ctx, N, X, y, NULL, NULL, &psi0, psi1, psi2, &sy2));
    CK(sgp_kuu_factor(ctx, 1e-6, NULL));
    CK(sgp_posterior_v(ctx, xi0, Lam0, 25.0, mu, Sig, Uv));
    CK(sgp_w_terms(ctx, mu, Uv, &s1, &s2));
    printf("%s\n", sgp_version());
    printf("N %d M %d D %d\n", N, M, D);
    printf("psi0 %.17g\nsum_y2 %.17g\nsumI1 %.17g\nsumI2 %.17g\n", psi0, sy2, s1, s2);
    printf("psi1"); for (int m = 0; m < M; ++m) printf(" %.17g", psi1[m]); printf("\n");
    printf("psi2_diag"); for (int m = 0; m < M; ++m) printf(" %.17g", psi2[(size_t)m * M + m]); printf("\n");
    printf("mu"); for (int m = 0; m < M; ++m) printf(" %.17g", mu[m]); printf("\n");
    printf("Uv_diag"); for (int m = 0; m < M; ++m) printf(" %.17g", Uv[(size_t)m * M + m]); printf("\n");
    sgp_destroy(ctx);
    free(X); free(y); free(Z); free(ell); free(psi1); free(psi2); free(xi0); free(Lam0); free(mu); free(Sig); free(Uv);
    return 0;
}
