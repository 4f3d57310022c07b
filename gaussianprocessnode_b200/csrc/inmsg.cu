// Backward message towards the INPUT of the MultiSGP node, for all nodes of a GPSSM / GPLVM chain at once.
//
// Reference: `@rule MultiSGP(:in, Marginalisation)` (GPnode/MultiSGPnode.jl:162-185, 187-211) returns, per node, the closure
//     log_backwardmess(x) = -1/2 tr(W) (Psi0(x) - sum(Kuu^-1 .* Psi2(x))) + sum(sumdiagV .* Psi1(x)) - 1/2 sum(Psi2(x) .* sumRvblk_W)
// with Psi0(x) = k(x,x), Psi1(x) = k_u(x), Psi2(x) = k_u(x) k_u(x)', sumdiagV = sum_d mu_v^(d) (W mu_y)_d, sumRvblk_W = sum_ij W_ij R_v^{ij};
// the third variant (:213-236) minimises its negative with LBFGS and returns the Laplace approximation (Hessian at the mode).
// The closure is evaluated by ReactiveMP at the cubature points of the forward message (the `prod` override, :38-45): 2d+1 points per
// node and VMP iteration, each an M-vector kernel column, an M x M outer product and two M x M Frobenius products in the reference.
//
// Batched: with k_p = k_u(x_p), A = S - tr(W) Kuu^-1 (S = sumRvblk_W, shared by all nodes) and g_n = Mv (W mu_y,n) = Mv r_n,
//     f_p = -1/2 tr(W) k(x_p,x_p) + g_n' k_p - 1/2 k_p' A k_p
//     grad f_p = J_p' (g_n - A k_p),                    J_p = d k_p / d x   (M x d)
//     hess f_p = sum_m (g_n - A k_p)_m  d2 k_pm / dx dx' - J_p' A J_p
// K = [k_p] (M x NP) and, for the derivatives, the d matrices J_j = [d k_p / d x_j]; A K and A J_j are tile GEMMs on the FP64 tensor
// pipe (sgp_gemm), the contractions one warp per point.  Analytic derivatives: SE-ARD only.
#include "sgp_internal.cuh"
#include <cmath>

int sgp_gemm(sgp_ctx* ctx, int opA, int opB, int m, int n, int k, double alpha, const double* A, int lda, const double* B, int ldb,
             double beta, double* C, int ldc, int lower_only);

namespace {

// K[m + p*M] = k(z_m, x_p);  J[j][m + p*M] = dk/dx_j = -k (x_j - z_mj) / ell_j^2   (SE only, nder = d or 0)
__global__ void in_kmat_kernel(const double* __restrict__ Xp, const double* __restrict__ Z, double* __restrict__ K, double* __restrict__ J,
                               long long NP, int M, int D, int kind, double variance, const double* __restrict__ ell_inv, int nder) {
    const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (e >= NP * M) return;
    const long long p = e / M;
    const int m = (int)(e - p * M);
    double r2 = 0.0, dlt[SGP_MAX_D];
    for (int d = 0; d < D; ++d) {
        const double t = (Xp[p * D + d] - Z[(size_t)m * D + d]) * ell_inv[d];
        dlt[d] = t;
        r2 = fma(t, t, r2);
    }
    double k;
    if (kind == SGP_KERNEL_SE) k = variance * exp(-0.5 * r2);
    else if (kind == SGP_KERNEL_MATERN32) { const double s = sqrt(3.0 * r2); k = variance * (1.0 + s) * exp(-s); }
    else { const double s = sqrt(5.0 * r2); k = variance * (1.0 + s + s * s / 3.0) * exp(-s); }
    K[e] = k;
    for (int j = 0; j < nder; ++j) J[(size_t)j * NP * M + e] = -k * dlt[j] * ell_inv[j];
}

__global__ void in_amat_kernel(double* __restrict__ A, const double* __restrict__ S, const double* __restrict__ Kinv, double trW, size_t n) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i < n) A[i] = S[i] - trW * Kinv[i];
}

// one warp per point p of node n = p / P:  f, grad (d), hess (d x d, column-major, symmetric)
__global__ void in_reduce_kernel(const double* __restrict__ Xp, const double* __restrict__ Z, const double* __restrict__ K, const double* __restrict__ Q,
                                 const double* __restrict__ J, const double* __restrict__ QJ, const double* __restrict__ Mv, const double* __restrict__ R,
                                 double* __restrict__ f, double* __restrict__ grad, double* __restrict__ hess, long long NP, int P, int M, int D, int Dout,
                                 double c0, const double* __restrict__ ell_inv, int want_hess) {
    const long long p = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (p >= NP) return;
    const long long n = p / P;
    double fv = 0.0, gv[SGP_MAX_D];
    for (int j = 0; j < D; ++j) gv[j] = 0.0;
    const bool der = grad != nullptr || hess != nullptr;
    for (int m = lane; m < M; m += 32) {
        double gm = 0.0;
        for (int d = 0; d < Dout; ++d) gm = fma(R ? R[n * Dout + d] : 1.0, Mv[(size_t)m + (size_t)d * M], gm);
        const double k = K[p * M + m], q = Q[p * M + m];
        fv = fma(k, gm - 0.5 * q, fv);
        if (der)
            for (int j = 0; j < D; ++j) gv[j] = fma(gm - q, J[(size_t)j * NP * M + p * M + m], gv[j]);
    }
    for (int o = 16; o > 0; o >>= 1) fv += __shfl_xor_sync(0xffffffffu, fv, o);
    if (lane == 0) f[p] = c0 + fv;
    if (der) {
        for (int j = 0; j < D; ++j) {
            double v = gv[j];
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0 && grad) grad[p * D + j] = v;
        }
    }
    if (hess && want_hess) {
        for (int i = 0; i < D; ++i)
            for (int j = 0; j <= i; ++j) {
                double h = 0.0;
                for (int m = lane; m < M; m += 32) {
                    double gm = 0.0;
                    for (int d = 0; d < Dout; ++d) gm = fma(R ? R[n * Dout + d] : 1.0, Mv[(size_t)m + (size_t)d * M], gm);
                    const double k = K[p * M + m], q = Q[p * M + m];
                    const double di = (Xp[p * D + i] - Z[(size_t)m * D + i]) * ell_inv[i] * ell_inv[i];
                    const double dj = (Xp[p * D + j] - Z[(size_t)m * D + j]) * ell_inv[j] * ell_inv[j];
                    const double d2k = k * (di * dj - (i == j ? ell_inv[i] * ell_inv[i] : 0.0));
                    h = fma(gm - q, d2k, h);
                    h = fma(-J[(size_t)i * NP * M + p * M + m], QJ[(size_t)j * NP * M + p * M + m], h);
                }
                for (int o = 16; o > 0; o >>= 1) h += __shfl_xor_sync(0xffffffffu, h, o);
                if (lane == 0) { hess[p * D * D + i + j * D] = h; hess[p * D * D + j + i * D] = h; }
            }
    }
}

// one warp per node: sums over the node's S sigma points of  w k(x,x),  w k' Kinv k,  w k' mu_v,  w |Uv k|^2
__global__ void node_terms_kernel(const double* __restrict__ K, const double* __restrict__ Q, const double* __restrict__ T, const double* __restrict__ wv,
                                  const double* __restrict__ mu, double* __restrict__ out, long long N, int S, int M, double variance) {
    const long long n = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (n >= N) return;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    for (int s = 0; s < S; ++s) {
        const long long p = n * S + s;
        const double w = wv[p];
        double q = 0.0, l = 0.0, t2 = 0.0;
        for (int m = lane; m < M; m += 32) {
            const double k = K[p * M + m], t = T[p * M + m];
            q = fma(k, Q[p * M + m], q);
            l = fma(k, mu[m], l);
            t2 = fma(t, t, t2);
        }
        a0 = fma(w, variance, a0); a1 = fma(w, q, a1); a2 = fma(w, l, a2); a3 = fma(w, t2, a3);
    }
    for (int o = 16; o > 0; o >>= 1) {
        a1 += __shfl_xor_sync(0xffffffffu, a1, o); a2 += __shfl_xor_sync(0xffffffffu, a2, o); a3 += __shfl_xor_sync(0xffffffffu, a3, o);
    }
    if (lane == 0) { out[n] = a0; out[N + n] = a1; out[2 * N + n] = a2; out[3 * N + n] = a3; }
}
// out[0] = trace of the M x M matrix A, out[1] = |B|_F^2   (single block, fixed order)
__global__ void trace_frob_kernel(const double* __restrict__ A, const double* __restrict__ B, int M, double* __restrict__ out) {
    __shared__ double s0[256], s1[256];
    double t = 0.0, f = 0.0;
    for (int i = threadIdx.x; i < M; i += 256) t += A[(size_t)i * M + i];
    for (size_t i = threadIdx.x; i < (size_t)M * M; i += 256) f = fma(B[i], B[i], f);
    s0[threadIdx.x] = t; s1[threadIdx.x] = f;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) { if ((int)threadIdx.x < o) { s0[threadIdx.x] += s0[threadIdx.x + o]; s1[threadIdx.x] += s1[threadIdx.x + o]; } __syncthreads(); }
    if (threadIdx.x == 0) { out[0] = s0[0]; out[1] = s1[0]; }
}

}  // namespace

extern "C" int sgp_uncertain_node_terms(sgp_ctx* ctx, const double* mu_v, const double* Uv, double* psi0_n, double* q_kinv_n, double* lin_n, double* q_rv_n,
                                        double* tr_kinv, double* frob_uv) {
    if (!ctx) return SGP_ERR_ARG;
    if (ctx->sp_N <= 0 || ctx->sp_S <= 0) SGP_FAIL(ctx, SGP_ERR_ARG, "uncertain_node_terms: needs the sigma-point cloud of a preceding sgp_sweep_psi_uncertain (methods 0-2)");
    if (!ctx->have_kuu || !ctx->Kinv_dev) SGP_FAIL(ctx, SGP_ERR_ARG, "uncertain_node_terms: sgp_kuu_factor first");
    if (!mu_v || !Uv) SGP_FAIL(ctx, SGP_ERR_ARG, "uncertain_node_terms: mu_v and Uv required");
    SGP_CUDA(ctx, cudaSetDevice(ctx->dev));
    const int M = ctx->M, D = ctx->D, S = ctx->sp_S;
    const long long N = ctx->sp_N, NP = N * S;
    const size_t MM = (size_t)M * M, KM = (size_t)NP * M;
    const size_t need = MM + M + SGP_MAX_D + 3 * KM + 4 * (size_t)N + 64;
    int rc = sgp_ensure(ctx, &ctx->in_dev, &ctx->in_cap, need); if (rc) return rc;
    double* Uvd = ctx->in_dev; double* mud = Uvd + MM; double* elld = mud + M; double* K = elld + SGP_MAX_D; double* Q = K + KM; double* T = Q + KM;
    double* out = T + KM; double* sc = out + 4 * (size_t)N;
    double ell_inv[SGP_MAX_D] = {0};
    for (int d = 0; d < D; ++d) ell_inv[d] = 1.0 / ctx->ell[d];
    cudaStream_t st = ctx->stream;
    SGP_CUDA(ctx, cudaMemcpyAsync(Uvd, Uv, MM * sizeof(double), cudaMemcpyHostToDevice, st));
    SGP_CUDA(ctx, cudaMemcpyAsync(mud, mu_v, (size_t)M * sizeof(double), cudaMemcpyHostToDevice, st));
    SGP_CUDA(ctx, cudaMemcpyAsync(elld, ell_inv, sizeof ell_inv, cudaMemcpyHostToDevice, st));
    in_kmat_kernel<<<(unsigned)((KM + 255) / 256), 256, 0, st>>>(ctx->sp_X_dev, ctx->Z_dev, K, nullptr, NP, M, D, ctx->kind, ctx->variance, elld, 0);
    SGP_CUDA(ctx, cudaGetLastError());
    rc = sgp_gemm(ctx, 0, 0, M, (int)NP, M, 1.0, ctx->Kinv_dev, M, K, M, 0.0, Q, M, 0); if (rc) return rc;
    rc = sgp_gemm(ctx, 0, 0, M, (int)NP, M, 1.0, Uvd, M, K, M, 0.0, T, M, 0); if (rc) return rc;
    node_terms_kernel<<<(unsigned)((N * 32 + 255) / 256), 256, 0, st>>>(K, Q, T, ctx->sp_w_dev, mud, out, N, S, M, ctx->variance);
    trace_frob_kernel<<<1, 256, 0, st>>>(ctx->Kinv_dev, Uvd, M, sc);
    SGP_CUDA(ctx, cudaGetLastError());
    if (psi0_n) SGP_CUDA(ctx, cudaMemcpyAsync(psi0_n, out, (size_t)N * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (q_kinv_n) SGP_CUDA(ctx, cudaMemcpyAsync(q_kinv_n, out + N, (size_t)N * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (lin_n) SGP_CUDA(ctx, cudaMemcpyAsync(lin_n, out + 2 * N, (size_t)N * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (q_rv_n) SGP_CUDA(ctx, cudaMemcpyAsync(q_rv_n, out + 3 * N, (size_t)N * sizeof(double), cudaMemcpyDeviceToHost, st));
    double sch[2] = {0.0, 0.0};
    SGP_CUDA(ctx, cudaMemcpyAsync(sch, sc, sizeof sch, cudaMemcpyDeviceToHost, st));
    SGP_CUDA(ctx, cudaStreamSynchronize(st));
    if (tr_kinv) *tr_kinv = sch[0];
    if (frob_uv) *frob_uv = sch[1];
    return SGP_OK;
}

extern "C" int sgp_in_logmessage(sgp_ctx* ctx, int64_t N, int P, const double* Xp, int D_out, const double* R, const double* Mv, const double* S,
                                 double trW, double* f, double* grad, double* hess) {
    if (!ctx) return SGP_ERR_ARG;
    if (!ctx->have_kernel || !ctx->have_Z) SGP_FAIL(ctx, SGP_ERR_ARG, "in_logmessage: set_kernel and set_inducing first");
    if (!ctx->have_kuu || !ctx->Kinv_dev) SGP_FAIL(ctx, SGP_ERR_ARG, "in_logmessage: sgp_kuu_factor first (K_uu^-1 is part of the message)");
    if (N < 1 || P < 1 || !Xp || D_out < 1 || !Mv || !S || !f) SGP_FAIL(ctx, SGP_ERR_ARG, "in_logmessage: N, P >= 1 and Xp, Mv, S, f required");
    const bool der = grad != nullptr || hess != nullptr;
    if (der && ctx->kind != SGP_KERNEL_SE) SGP_FAIL(ctx, SGP_ERR_UNSUPPORTED, "in_logmessage: analytic gradient / Hessian for the SE-ARD kernel only");
    SGP_CUDA(ctx, cudaSetDevice(ctx->dev));
    const int M = ctx->M, D = ctx->D;
    const long long NP = (long long)N * P;
    const size_t MM = (size_t)M * M, KM = (size_t)NP * M;
    const int nder = der ? D : 0;
    // scratch: A | Mv | S | ell_inv | Xp | R | K | Q | J[d] | QJ[d] | f | grad | hess
    const size_t need = MM * 2 + (size_t)M * D_out + SGP_MAX_D + (size_t)NP * D + (size_t)N * D_out + KM * (2 + 2 * (size_t)nder) + (size_t)NP * (1 + D + (size_t)D * D) + 64;
    int rc = sgp_ensure(ctx, &ctx->in_dev, &ctx->in_cap, need); if (rc) return rc;
    double* A = ctx->in_dev; double* Sd = A + MM; double* Mvd = Sd + MM; double* elld = Mvd + (size_t)M * D_out;
    double* Xd = elld + SGP_MAX_D; double* Rd = Xd + (size_t)NP * D; double* K = Rd + (size_t)N * D_out; double* Q = K + KM;
    double* J = Q + KM; double* QJ = J + KM * nder; double* fd = QJ + KM * nder; double* gd = fd + NP; double* hd = gd + (size_t)NP * D;
    double ell_inv[SGP_MAX_D] = {0};
    for (int d = 0; d < D; ++d) ell_inv[d] = 1.0 / ctx->ell[d];
    cudaStream_t st = ctx->stream;
    SGP_CUDA(ctx, cudaMemcpyAsync(Sd, S, MM * sizeof(double), cudaMemcpyHostToDevice, st));
    SGP_CUDA(ctx, cudaMemcpyAsync(Mvd, Mv, (size_t)M * D_out * sizeof(double), cudaMemcpyHostToDevice, st));
    SGP_CUDA(ctx, cudaMemcpyAsync(elld, ell_inv, sizeof ell_inv, cudaMemcpyHostToDevice, st));
    SGP_CUDA(ctx, cudaMemcpyAsync(Xd, Xp, (size_t)NP * D * sizeof(double), cudaMemcpyHostToDevice, st));
    if (R) SGP_CUDA(ctx, cudaMemcpyAsync(Rd, R, (size_t)N * D_out * sizeof(double), cudaMemcpyHostToDevice, st));
    in_amat_kernel<<<(unsigned)((MM + 255) / 256), 256, 0, st>>>(A, Sd, ctx->Kinv_dev, trW, MM);
    in_kmat_kernel<<<(unsigned)((KM + 255) / 256), 256, 0, st>>>(Xd, ctx->Z_dev, K, J, NP, M, D, ctx->kind, ctx->variance, elld, nder);
    SGP_CUDA(ctx, cudaGetLastError());
    rc = sgp_gemm(ctx, 0, 0, M, (int)NP, M, 1.0, A, M, K, M, 0.0, Q, M, 0); if (rc) return rc;
    if (hess)
        for (int j = 0; j < D; ++j) { rc = sgp_gemm(ctx, 0, 0, M, (int)NP, M, 1.0, A, M, J + KM * j, M, 0.0, QJ + KM * j, M, 0); if (rc) return rc; }
    // k(x,x): sigma^2 for all three stationary kernels
    const double c0 = -0.5 * trW * ctx->variance;
    in_reduce_kernel<<<(unsigned)((NP * 32 + 255) / 256), 256, 0, st>>>(Xd, ctx->Z_dev, K, Q, J, QJ, Mvd, R ? Rd : nullptr, fd, grad || hess ? gd : nullptr,
                                                                         hess ? hd : nullptr, NP, P, M, D, D_out, c0, elld, hess ? 1 : 0);
    SGP_CUDA(ctx, cudaGetLastError());
    SGP_CUDA(ctx, cudaMemcpyAsync(f, fd, (size_t)NP * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (grad) SGP_CUDA(ctx, cudaMemcpyAsync(grad, gd, (size_t)NP * D * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (hess) SGP_CUDA(ctx, cudaMemcpyAsync(hess, hd, (size_t)NP * D * D * sizeof(double), cudaMemcpyDeviceToHost, st));
    SGP_CUDA(ctx, cudaStreamSynchronize(st));
    return SGP_OK;
}
