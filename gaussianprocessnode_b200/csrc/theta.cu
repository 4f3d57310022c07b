// The theta step's collapsed objective and its exact gradient.
//
// Replaces `neg_log_backwardmess_fast` + `ForwardDiff.gradient!` (helper_functions/derivative_helper.jl:23-39, 55-67; called
// once per mini-batch in experiments/regression_kin40k.ipynb:214-222), where a per-point loop does `Lu \ k_n`, `Uv * k_n`
// in dual-number arithmetic:
//     F(theta) = sum_n [ w/2 k_nn - w/2 |L_u^-1 k_n|^2 + w/2 |U_v k_n|^2 - w y_n v' k_n ]
//              = w/2 [ Psi0 - <K_uu^-1, Psi2> + <R_v, Psi2> ] - w v' Psi1,          R_v = U_v' U_v
// The value comes from the statistics of the fused sweep.  The gradient is analytic (what forward-mode AD evaluates, in closed
// form): with A = R_v - K_uu^-1, B = K_uu^-1 Psi2 K_uu^-1, and d k(x,z) / d ell_d = h(r) (x_d - z_d)^2 / ell_d^3
// (h = k for SE, sigma^2 3 e^{-sqrt3 r} for Matern-3/2, sigma^2 5/3 (1 + sqrt5 r) e^{-sqrt5 r} for Matern-5/2),
//     dF/d ell_d = [ sum_n sum_m h_nm c_nmd (w (A k_n)_m - w y_n v_m) + w/2 sum_mm' B_mm' h_mm' c_mm'd ] / ell_d^3
//     dF/d sigma^2 = [ w/2 Psi0 - w/2 <K_uu^-1, Psi2> + w <R_v, Psi2> - w v' Psi1 - w/2 jitter tr B ] / sigma^2
// A k_n for a chunk of points is one DMMA GEMM against the materialised K_uf chunk; the contraction is an FP64-ALU-bound
// elementwise pass with fixed-order partial sums (deterministic).  All M x M algebra reuses dense.cu.
#include "sgp_internal.cuh"
#include <cmath>
#include <algorithm>
#include <vector>
#include <utility>
#include <cstdio>
#include <cstdlib>

int sgp_gemm(sgp_ctx* ctx, int opA, int opB, int m, int n, int k, double alpha, const double* A, int lda, const double* B, int ldb,
             double beta, double* C, int ldc, int lower_only);

namespace {

struct KParams {
    int kind, D, M;
    double variance;
    double ell_inv[SGP_MAX_D];
};

// DP = D rounded up to 4 / 8 / 16: the unrolled loops cover DP dimensions only
template <int DP = SGP_MAX_D>
__device__ __forceinline__ double r2_of(const double* __restrict__ x, const double* __restrict__ z, const KParams& kp, double* c) {
    double r2 = 0.0;
#pragma unroll                                                    // (static indices: c stays in registers)
    for (int d = 0; d < DP; ++d) {
        const double t = d < kp.D ? x[d] - z[d] : 0.0;
        c[d] = t * t;                                             // (x_d - z_d)^2, unscaled
        r2 = fma(c[d], kp.ell_inv[d] * kp.ell_inv[d], r2);
    }
    return r2;
}
__device__ __forceinline__ double k_of(const KParams& kp, double r2) {
    if (kp.kind == SGP_KERNEL_SE) return kp.variance * exp(-0.5 * r2);
    const double s = sqrt((kp.kind == SGP_KERNEL_MATERN32 ? 3.0 : 5.0) * r2);
    return kp.kind == SGP_KERNEL_MATERN32 ? kp.variance * (1.0 + s) * exp(-s) : kp.variance * (1.0 + s + s * s / 3.0) * exp(-s);
}
// d k / d ell_d = h * (x_d - z_d)^2 / ell_d^3
__device__ __forceinline__ double h_of(const KParams& kp, double r2) {
    if (kp.kind == SGP_KERNEL_SE) return kp.variance * exp(-0.5 * r2);
    if (kp.kind == SGP_KERNEL_MATERN32) return kp.variance * 3.0 * exp(-sqrt(3.0 * r2));
    const double s = sqrt(5.0 * r2);
    return kp.variance * (5.0 / 3.0) * (1.0 + s) * exp(-s);
}

// K[m + j*M] = k(z_m, x_{n0+j})
__global__ void kuf_chunk_kernel(const double* __restrict__ X, const double* __restrict__ Z, double* __restrict__ K, long long n0, int nc, KParams kp) {
    const size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (e >= (size_t)kp.M * nc) return;
    const int m = (int)(e % kp.M); const long long j = (long long)(e / kp.M);
    double c[SGP_MAX_D];
    K[e] = k_of(kp, r2_of(X + (n0 + j) * kp.D, Z + (size_t)m * kp.D, kp, c));
}

// Block reduction of acc[0 .. SGP_MAX_D] (slot SGP_MAX_D: an extra scalar) into partial[block][.]; the LAST block to arrive (ticket) adds the
// partials of all blocks in fixed order -- deterministic -- and accumulates total[d] += scale * sum (extra: *extra_out += sum, if given).
constexpr int kPS = SGP_MAX_D + 1;
__device__ __forceinline__ void reduce_and_finish(const double (&acc)[kPS], double* __restrict__ partial, unsigned* __restrict__ ticket, double scale,
                                                  double* __restrict__ total, double* __restrict__ extra_out) {
    __shared__ double s[8][kPS];
    __shared__ int s_last;
    const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
#pragma unroll
    for (int d = 0; d < kPS; ++d) {
        double a = acc[d];
        for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
        if (lane == 0) s[wp][d] = a;
    }
    __syncthreads();
    if (threadIdx.x < kPS) {
        double a = 0.0;
        for (int q = 0; q < 8; ++q) a += s[q][threadIdx.x];
        partial[(size_t)threadIdx.x * gridDim.x + blockIdx.x] = a;          // [d][block]: the finish reads along the blocks
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const bool last = atomicAdd(ticket, 1u) == gridDim.x - 1;
        if (last) *ticket = 0u;
        s_last = last ? 1 : 0;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    for (int d = wp; d < kPS; d += 8) {
        double a = 0.0;
#pragma unroll 4
        for (int b = lane; b < (int)gridDim.x; b += 32) a += __ldcg(partial + (size_t)d * gridDim.x + b);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
        if (lane == 0) {
            if (d < SGP_MAX_D) total[d] += scale * a;
            else if (extra_out) *extra_out += a;
        }
    }
}

// total[d] += sum over the chunk's (m, j) of h_mj c_mjd (w G_mj - w y_j v_m)
template <int DP>
__global__ void __launch_bounds__(256) grad_contract_kernel(const double* __restrict__ X, const double* __restrict__ y, const double* __restrict__ Z,
                                                            const double* __restrict__ G, const double* __restrict__ v, double w, long long n0,
                                                            int nc, KParams kp, double* __restrict__ partial, unsigned* __restrict__ ticket,
                                                            double* __restrict__ total_out) {
    double acc[kPS];
#pragma unroll
    for (int d = 0; d < kPS; ++d) acc[d] = 0.0;
    const size_t total = (size_t)kp.M * nc;
    for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
        const int m = (int)(e % kp.M); const long long j = (long long)(e / kp.M);
        double c[DP];
        const double r2 = r2_of<DP>(X + (n0 + j) * kp.D, Z + (size_t)m * kp.D, kp, c);
        const double f = h_of(kp, r2) * (w * G[e] - w * y[n0 + j] * v[m]);
#pragma unroll
        for (int d = 0; d < DP; ++d) acc[d] = fma(f, c[d], acc[d]);       // (c[d] = 0 beyond D)
    }
    reduce_and_finish(acc, partial, ticket, 1.0, total_out, nullptr);
}

// total[d] += scale * sum over (m, m') of B_mm' h_mm' c_mm'd   (K_uu part of the gradient);  *trace_out += tr B
template <int DP>
__global__ void __launch_bounds__(256) kuu_contract_kernel(const double* __restrict__ Z, const double* __restrict__ B, KParams kp, double* __restrict__ partial,
                                                           unsigned* __restrict__ ticket, double scale, double* __restrict__ total_out,
                                                           double* __restrict__ trace_out) {
    double acc[kPS];
#pragma unroll
    for (int d = 0; d < kPS; ++d) acc[d] = 0.0;
    const size_t total = (size_t)kp.M * kp.M;
    for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
        const int m = (int)(e % kp.M), m2 = (int)(e / kp.M);
        double c[DP];
        const double r2 = r2_of<DP>(Z + (size_t)m * kp.D, Z + (size_t)m2 * kp.D, kp, c);
        const double b = B[e];
        if (m == m2) acc[SGP_MAX_D] += b;
        const double f = h_of(kp, r2) * b;
#pragma unroll
        for (int d = 0; d < DP; ++d) acc[d] = fma(f, c[d], acc[d]);
    }
    reduce_and_finish(acc, partial, ticket, scale, total_out, trace_out);
}

// res[8 .. 12] = the four scalars of the statistics; res[12] = the pivot record of the factorisations: ONE read-back for the whole step
__global__ void gather_kernel(const double* __restrict__ scal, const int* __restrict__ info, double* __restrict__ res) {
    if (threadIdx.x < 4) res[8 + threadIdx.x] = scal[threadIdx.x];
    if (threadIdx.x == 4) res[12] = (double)*info;
}

inline unsigned nb(size_t n, int t = 256) { return (unsigned)((n + t - 1) / t); }

// A = (R ? R : Sig + mu mu') - Kinv, and Rv (optional) = that first term.  R: only its lower triangle is read (Rv may be R itself: the
// mirror is written into the part nobody reads).
__global__ void amat_kernel(double* __restrict__ A, double* Rv, const double* R, const double* __restrict__ Sig, const double* __restrict__ mu,
                            const double* __restrict__ Kinv, int M) {
    const size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (e >= (size_t)M * M) return;
    const size_t i = e % M, j = e / M;
    const double r = R ? R[i >= j ? e : j + i * M] : fma(mu[i], mu[j], Sig[e]);
    if (Rv) Rv[e] = r;
    A[e] = r - Kinv[e];
}

}  // namespace

extern "C" int sgp_theta_objective(sgp_ctx* ctx, const double* mu_v, const double* Uv, double w, double jitter, double* value,
                                   double* dvariance, double* dlengthscale) {
    if (!ctx) return SGP_ERR_ARG;
    if (!ctx->have_kernel || !ctx->have_Z || !ctx->have_data) SGP_FAIL(ctx, SGP_ERR_ARG, "theta_objective: set_kernel, set_inducing and set_data first");
    if ((mu_v == nullptr) != (Uv == nullptr)) SGP_FAIL(ctx, SGP_ERR_ARG, "theta_objective: pass both mu_v and Uv, or neither (resident posterior)");
    if (!mu_v && !(sgp_resident_mu(ctx) && sgp_resident_sigma(ctx))) SGP_FAIL(ctx, SGP_ERR_ARG, "theta_objective: no resident posterior (sgp_posterior_v first)");
    if (ctx->have_w) SGP_FAIL(ctx, SGP_ERR_UNSUPPORTED, "theta_objective: per-point weights are not part of the reference objective");
    SGP_RANGE("sgp_theta_objective");
    SGP_CUDA(ctx, cudaSetDevice(ctx->dev));
    const int M = ctx->M, D = ctx->D;
    const size_t MM = (size_t)M * M;
    const int64_t N = ctx->N;
    const bool want_grad = dvariance != nullptr || dlengthscale != nullptr;

    // The objective is linear in the statistics of the resident data at the current kernel: they are reused when the last sweep produced exactly
    // those, otherwise the regular sweep runs (with a communicator: summed over the ranks, so that the resident statistics stay what every
    // other call expects).  Only the data part of the gradient below is rank-local and summed at the end.
    int rc = SGP_OK;
    // tuning aid (SGP_THETA_TIMING=1): device time of every stage on stdout
    static const bool timing = std::getenv("SGP_THETA_TIMING") != nullptr;
    std::vector<std::pair<const char*, cudaEvent_t>> marks;
    auto mark = [&](const char* name) {
        if (!timing) return;
        cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, ctx->stream); marks.emplace_back(name, e);
    };
    mark("start");
    if (!(ctx->have_stats && ctx->stats_of_data && ctx->Dout == 1)) { rc = sgp_sweep_resident(ctx, false); if (rc) return rc; }
    mark("sweep");
    // (enqueued only: the whole step has ONE host synchronisation, at the end, where the pivot record is checked too)
    const bool factored_here = !(ctx->have_kuu && ctx->kuu_jitter == jitter);
    if (factored_here) { rc = sgp_kuu_factor_enqueue(ctx, jitter); if (rc) return rc; }
    mark("K_uu job");

    const int nc_max = (int)std::min<int64_t>(N, std::max<int64_t>(1024, (int64_t)(16u << 20) / M));     // K and G chunks: 2 x 128 MB at most
    const int cblocks = 2 * ctx->num_sms;
    // scratch: [Rv | A | B | T] (M x M each) | v (M) | res (64) | totals (2 x SGP_MAX_D) | K chunk | G chunk | block partials
    size_t need = 4 * MM + (size_t)M + 64 + 2 * SGP_MAX_D + (want_grad ? 2 * (size_t)M * nc_max + (size_t)cblocks * kPS : 0);
    rc = sgp_ensure(ctx, &ctx->theta_dev, &ctx->theta_cap, need); if (rc) return rc;
    double* Rv = ctx->theta_dev; double* A = Rv + MM; double* B = A + MM; double* T = B + MM;
    double* vdev = T + MM; double* res = vdev + M; double* total = res + 64; double* total_data = total + SGP_MAX_D; double* Kc = total_data + SGP_MAX_D;
    double* Gc = Kc + (size_t)M * nc_max; double* partial = Gc + (size_t)M * nc_max;
    double* psi2 = ctx->stats_dev; double* psi1 = psi2 + MM; double* scal = psi1 + M;
    const double* Kinv = ctx->Kinv_dev;

    // R_v = Uv' Uv (host posterior) or Sigma_v + mu_v mu_v' (resident posterior: no Cholesky factor needed); A = R_v - K_uu^-1
    const double* v = nullptr;
    if (mu_v) {
        // (on the copy stream: the upload runs beside the sweep and the K_uu job enqueued above; T and vdev are free -- every earlier call
        //  that used them has been synchronised by its own final read-back)
        SGP_CUDA(ctx, cudaMemcpyAsync(T, Uv, MM * sizeof(double), cudaMemcpyHostToDevice, ctx->stream2));
        SGP_CUDA(ctx, cudaMemcpyAsync(vdev, mu_v, (size_t)M * sizeof(double), cudaMemcpyHostToDevice, ctx->stream2));
        SGP_CUDA(ctx, cudaEventRecord(ctx->ev_copy, ctx->stream2));
        SGP_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_copy, 0));
        rc = sgp_gemm(ctx, 1, 0, M, M, M, 1.0, T, M, T, M, 0.0, Rv, M, 2); if (rc) return rc;      // lower triangle, triangular K range
        amat_kernel<<<nb(MM), 256, 0, ctx->stream>>>(A, Rv, Rv, nullptr, nullptr, Kinv, M);
        v = vdev;
    } else {
        v = sgp_resident_mu(ctx);
        amat_kernel<<<nb(MM), 256, 0, ctx->stream>>>(A, Rv, nullptr, sgp_resident_sigma(ctx), v, Kinv, M);
    }
    mark("R_v, A");
    // scalars in one pass: <Kinv, Psi2>, <Rv, Psi2>, v' Psi1
    rc = sgp_wterms_reduce(ctx, Kinv, psi2, Rv, nullptr, v, psi1, M, res); if (rc) return rc;
    mark("scalars");
    SGP_CUDA(ctx, cudaMemsetAsync(res + 4, 0, (60 + 2 * SGP_MAX_D) * sizeof(double), ctx->stream));      // tr B, the ticket word and both totals (contiguous)
    unsigned* ticket = reinterpret_cast<unsigned*>(res + 5);

    if (want_grad) {
        KParams kp{}; kp.kind = ctx->kind; kp.D = D; kp.M = M; kp.variance = ctx->variance;
        for (int d = 0; d < D; ++d) kp.ell_inv[d] = 1.0 / ctx->ell[d];
        // B = Kinv Psi2 Kinv, tr B, K_uu part of the lengthscale gradient (from the rank-summed statistics: identical on every rank)
        rc = sgp_gemm(ctx, 0, 0, M, M, M, 1.0, Kinv, M, psi2, M, 0.0, T, M, 0); if (rc) return rc;
        // (B in full, not its lower triangle mirrored: with an ill-conditioned K_uu the rounding of (Kinv Psi2) Kinv is far from symmetric
        //  and the contraction below relies on the two halves averaging out, as the reference's does)
        rc = sgp_gemm(ctx, 0, 0, M, M, M, 1.0, T, M, Kinv, M, 0.0, B, M, 0); if (rc) return rc;
        mark("B = Kinv Psi2 Kinv");
        if (D <= 4) kuu_contract_kernel<4><<<cblocks, 256, 0, ctx->stream>>>(ctx->Z_dev, B, kp, partial, ticket, 0.5 * w, total, res + 4);
        else if (D <= 8) kuu_contract_kernel<8><<<cblocks, 256, 0, ctx->stream>>>(ctx->Z_dev, B, kp, partial, ticket, 0.5 * w, total, res + 4);
        else kuu_contract_kernel<16><<<cblocks, 256, 0, ctx->stream>>>(ctx->Z_dev, B, kp, partial, ticket, 0.5 * w, total, res + 4);
        mark("K_uu contraction");
        // data part (this rank's points), chunk by chunk: K chunk -> G = A K -> contraction
        for (int64_t n0 = 0; n0 < N; n0 += nc_max) {
            const int nc = (int)std::min<int64_t>(nc_max, N - n0);
            kuf_chunk_kernel<<<nb((size_t)M * nc), 256, 0, ctx->stream>>>(ctx->X_dev, ctx->Z_dev, Kc, n0, nc, kp);
            rc = sgp_gemm(ctx, 0, 0, M, nc, M, 1.0, A, M, Kc, M, 0.0, Gc, M, 0); if (rc) return rc;
            if (D <= 4) grad_contract_kernel<4><<<cblocks, 256, 0, ctx->stream>>>(ctx->X_dev, ctx->y_dev, ctx->Z_dev, Gc, v, w, n0, nc, kp, partial, ticket, total_data);
            else if (D <= 8) grad_contract_kernel<8><<<cblocks, 256, 0, ctx->stream>>>(ctx->X_dev, ctx->y_dev, ctx->Z_dev, Gc, v, w, n0, nc, kp, partial, ticket, total_data);
            else grad_contract_kernel<16><<<cblocks, 256, 0, ctx->stream>>>(ctx->X_dev, ctx->y_dev, ctx->Z_dev, Gc, v, w, n0, nc, kp, partial, ticket, total_data);
        }
        mark("data part (K chunk, G = A K, contraction)");
        if (ctx->comm) { rc = sgp_comm_allreduce(ctx, total_data, (size_t)SGP_MAX_D); if (rc) return rc; }     // the only rank-local part
    }
    SGP_CUDA(ctx, cudaGetLastError());
    // ONE read-back and ONE host synchronisation for the whole step: [res (64) | totals (2 x SGP_MAX_D)] are contiguous
    gather_kernel<<<1, 32, 0, ctx->stream>>>(scal, ctx->info_dev, res);
    SGP_CUDA(ctx, cudaGetLastError());
    double* hb = sgp_host_stage(ctx, 64 + 2 * SGP_MAX_D);
    if (!hb) SGP_FAIL(ctx, SGP_ERR_CUDA, "theta_objective: pinned staging buffer");
    SGP_CUDA(ctx, cudaMemcpyAsync(hb, res, (64 + 2 * SGP_MAX_D) * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    mark("read-back");
    SGP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (timing) {
        printf("theta step M=%d N=%lld:", M, (long long)N);
        for (size_t i = 1; i < marks.size(); ++i) { float ms = 0.f; cudaEventElapsedTime(&ms, marks[i - 1].second, marks[i].second); printf(" %s %.1f us |", marks[i].first, ms * 1e3f); }
        float ms = 0.f; cudaEventElapsedTime(&ms, marks.front().second, marks.back().second); printf(" total %.1f us\n", ms * 1e3f);
        for (auto& m : marks) cudaEventDestroy(m.second);
        fflush(stdout);
    }
    const double* h = hb; const double* sc = hb + 8; const double* tot = hb + 64;
    if (factored_here && hb[12] != 0.0) {
        ctx->have_kuu = false;
        char buf[160];
        snprintf(buf, sizeof buf, "theta_objective: Cholesky of K_uu met a non-positive pivot at row %d of %d", (int)hb[12], M);
        SGP_FAIL(ctx, SGP_ERR_NOT_PD, buf);
    }
    if (value) *value = 0.5 * w * (sc[0] - h[0] + h[1]) - w * h[2];
    if (dvariance) *dvariance = (0.5 * w * sc[0] - 0.5 * w * h[0] + w * h[1] - w * h[2] - 0.5 * w * jitter * h[4]) / ctx->variance;
    if (dlengthscale) for (int d = 0; d < D; ++d) dlengthscale[d] = (tot[d] + tot[SGP_MAX_D + d]) / (ctx->ell[d] * ctx->ell[d] * ctx->ell[d]);
    return SGP_OK;
}
