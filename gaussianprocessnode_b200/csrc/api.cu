// C ABI of libsgp (include/sgp.h): context, resident state, and the calls a ReactiveMP host makes in place of the
// per-point rules.  Host pointers in, host pointers out; all device work on the context's own stream.
#include "sgp_internal.cuh"
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <new>

int sgp_gemm(sgp_ctx* ctx, int opA, int opB, int m, int n, int k, double alpha, const double* A, int lda, const double* B, int ldb,
             double beta, double* C, int ldc, int lower_only);
int sgp_uncertain_sweep(sgp_ctx* ctx, int method, int p, int64_t N, const double* mean, const double* cov, int D_out, const double* R,
                        double* psi0, double* psi1, double* psi2, double* psi1_n);

namespace {

// out[n] = sum_m k(x_n, z_m) mu[m]   (one thread per test point, inducing rows broadcast from shared memory)
// prob (optional) = Phi(out[n] * inv_s): the Probit(:out) message of the :out message N(out[n], 1 / w_bar), inv_s = 1 / sqrt(1 + 1 / w_bar)
__global__ void predict_kernel(const double* __restrict__ Xt, const double* __restrict__ Z, const double* __restrict__ mu, double* __restrict__ out,
                               long long Nt, int M, int D, int kind, double variance, const double* __restrict__ ell_inv, double* __restrict__ prob,
                               double inv_s) {
    extern __shared__ double sh[];   // chunk of inducing rows: [MC][D] + mu[MC]
    const int MC = 256;
    long long n = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    double x[SGP_MAX_D];
    for (int d = 0; d < D; ++d) x[d] = n < Nt ? Xt[n * D + d] * ell_inv[d] : 0.0;
    double acc = 0.0;
    for (int m0 = 0; m0 < M; m0 += MC) {
        int mc = M - m0 < MC ? M - m0 : MC;
        __syncthreads();
        for (int e = threadIdx.x; e < mc * D; e += blockDim.x) sh[e] = Z[(size_t)m0 * D + e] * ell_inv[e % D];
        for (int e = threadIdx.x; e < mc; e += blockDim.x) sh[MC * D + e] = mu[m0 + e];
        __syncthreads();
        for (int m = 0; m < mc; ++m) {
            double r2 = 0.0;
            for (int d = 0; d < D; ++d) { double t = x[d] - sh[m * D + d]; r2 = fma(t, t, r2); }
            double k;
            if (kind == SGP_KERNEL_SE) k = exp(-0.5 * r2);
            else if (kind == SGP_KERNEL_MATERN32) { double s = sqrt(3.0 * r2); k = (1.0 + s) * exp(-s); }
            else { double s = sqrt(5.0 * r2); k = (1.0 + s + s * s / 3.0) * exp(-s); }
            acc = fma(k, sh[MC * D + m], acc);
        }
    }
    if (n < Nt) {
        out[n] = variance * acc;
        if (prob) prob[n] = 0.5 * erfc(-variance * acc * inv_s * 0.70710678118654752440);
    }
}

inline unsigned nblocks(size_t n, int t = 256) { return (unsigned)((n + t - 1) / t); }

int check(sgp_ctx* ctx) { return ctx ? SGP_OK : SGP_ERR_ARG; }

}  // namespace

// like sgp_ensure, zero-filled when (re)allocated: the Cholesky writes only the lower triangle of the diagonal-block inverses
int sgp_ensure_zero(sgp_ctx* ctx, double** p, size_t* cap, size_t need) {
    if (*cap >= need && *p) return SGP_OK;
    int rc = sgp_ensure(ctx, p, cap, need); if (rc) return rc;
    SGP_CUDA(ctx, cudaMemsetAsync(*p, 0, need * sizeof(double), ctx->stream));
    return SGP_OK;
}

int sgp_ensure(sgp_ctx* ctx, double** p, size_t* cap, size_t need) {
    if (*cap >= need && *p) return SGP_OK;
    if (*p) SGP_CUDA(ctx, cudaFree(*p));
    *p = nullptr; *cap = 0;
    SGP_CUDA(ctx, cudaMalloc((void**)p, need * sizeof(double)));
    *cap = need;
    return SGP_OK;
}

extern "C" {

const char* sgp_version(void) { return "libsgp 0.1 sm_100a"; }

int sgp_pinned_alloc(size_t bytes, void** ptr) {
    if (!ptr || bytes == 0) return SGP_ERR_ARG;
    *ptr = nullptr;
    return cudaHostAlloc(ptr, bytes, cudaHostAllocPortable) == cudaSuccess ? SGP_OK : SGP_ERR_CUDA;
}
void sgp_pinned_free(void* ptr) { if (ptr) cudaFreeHost(ptr); }

const char* sgp_last_error(const sgp_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int sgp_create(sgp_ctx** out, int device_id) {
    if (!out) return SGP_ERR_ARG;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device_id < 0 || device_id >= ndev) return SGP_ERR_CUDA;
    sgp_ctx* ctx = new (std::nothrow) sgp_ctx();
    if (!ctx) return SGP_ERR_CUDA;
    ctx->dev = device_id;
    auto fail = [&](const char*) { delete ctx; return SGP_ERR_CUDA; };
    if (cudaSetDevice(device_id) != cudaSuccess) return fail("setdevice");
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device_id) != cudaSuccess) return fail("props");
    if (prop.major < 10) { delete ctx; return SGP_ERR_UNSUPPORTED; }
    ctx->num_sms = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) return fail("stream");
    if (cudaStreamCreateWithFlags(&ctx->stream2, cudaStreamNonBlocking) != cudaSuccess) return fail("stream");
    if (cudaEventCreateWithFlags(&ctx->ev_copy, cudaEventDisableTiming) != cudaSuccess) return fail("event");
    for (auto& e : ctx->ev) if (cudaEventCreate(&e) != cudaSuccess) return fail("event");
    if (cudaMalloc((void**)&ctx->exptab_dev, SGP_EXP_TAB * sizeof(double)) != cudaSuccess) return fail("malloc");
    if (cudaMalloc((void**)&ctx->info_dev, 2 * sizeof(int)) != cudaSuccess) return fail("malloc");
    std::vector<double> tab(SGP_EXP_TAB);
    for (int j = 0; j < SGP_EXP_TAB; ++j) tab[j] = std::exp2((double)j / SGP_EXP_TAB);
    if (cudaMemcpy(ctx->exptab_dev, tab.data(), SGP_EXP_TAB * sizeof(double), cudaMemcpyHostToDevice) != cudaSuccess) return fail("memcpy");
    *out = ctx;
    return SGP_OK;
}

void sgp_destroy(sgp_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->dev);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    sgp_comm_destroy(ctx);
    if (ctx->own_data) { cudaFree(ctx->X_dev); cudaFree(ctx->y_dev); cudaFree(ctx->yv_dev); cudaFree(ctx->w_dev); }
    cudaFree(ctx->Z_dev); cudaFree(ctx->stats_dev); cudaFree(ctx->work_dev); cudaFree(ctx->zrec_dev); cudaFree(ctx->exptab_dev);
    cudaFree(ctx->dense_dev); cudaFree(ctx->info_dev); cudaFree(ctx->KuuL_dev); cudaFree(ctx->sp_X_dev); cudaFree(ctx->sp_w_dev);
    cudaFree(ctx->sp_y_dev); cudaFree(ctx->sweep_dbg_dev); cudaFree(ctx->theta_dev); cudaFree(ctx->Kinv_dev); cudaFree(ctx->kuu_dinv_dev);
    cudaFree(ctx->dinv_dev); cudaFree(ctx->post_dev); cudaFree(ctx->unc_dev); if (ctx->fetch_host) cudaFreeHost(ctx->fetch_host);
    cudaFree(ctx->kbuf_dev); cudaFree(ctx->sweep_flags_dev); cudaFree(ctx->flush_dev); cudaFree(ctx->in_dev); cudaFree(ctx->wt_dev); cudaFree(ctx->pred_dev);
    for (auto& e : ctx->ev) if (e) cudaEventDestroy(e);
    if (ctx->ev_copy) cudaEventDestroy(ctx->ev_copy);
    if (ctx->ready_dev) cudaFree(ctx->ready_dev);
    if (ctx->ready_host) cudaFreeHost(ctx->ready_host);
    if (ctx->packed_dev) cudaFree(ctx->packed_dev);
    if (ctx->p2plan_dev) cudaFree(ctx->p2plan_dev);
    if (ctx->stream2) cudaStreamDestroy(ctx->stream2);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

int sgp_set_kernel(sgp_ctx* ctx, int kind, int D, double variance, const double* lengthscale) {
    if (check(ctx)) return SGP_ERR_ARG;
    if (kind < 0 || kind > 2 || D < 1 || D > SGP_MAX_D || !lengthscale || !(variance > 0.0)) SGP_FAIL(ctx, SGP_ERR_ARG, "set_kernel: bad kind / D (1..16) / variance");
    for (int d = 0; d < D; ++d) if (!(lengthscale[d] > 0.0)) SGP_FAIL(ctx, SGP_ERR_ARG, "set_kernel: lengthscale must be positive");
    if (ctx->have_Z && D != ctx->D) { ctx->have_Z = false; }
    if (ctx->have_data && D != ctx->D) { ctx->N = 0; ctx->have_data = false; }
    ctx->kind = kind; ctx->D = D; ctx->variance = variance;
    for (int d = 0; d < D; ++d) ctx->ell[d] = lengthscale[d];
    ctx->have_kernel = true; ctx->have_kuu = false; ctx->have_stats = false;
    return SGP_OK;
}

int sgp_set_inducing(sgp_ctx* ctx, int M, const double* Z) {
    if (check(ctx)) return SGP_ERR_ARG;
    if (!ctx->have_kernel) SGP_FAIL(ctx, SGP_ERR_ARG, "set_inducing: set_kernel first (D is taken from it)");
    if (M < 1 || !Z) SGP_FAIL(ctx, SGP_ERR_ARG, "set_inducing: M >= 1 and Z required");
    SGP_CUDA(ctx, cudaSetDevice(ctx->dev));
    const int D = ctx->D;
    if (ctx->Z_dev) { SGP_CUDA(ctx, cudaFree(ctx->Z_dev)); ctx->Z_dev = nullptr; }
    SGP_CUDA(ctx, cudaMalloc((void**)&ctx->Z_dev, (size_t)M * D * sizeof(double)));
    SGP_CUDA(ctx, cudaMemcpyAsync(ctx->Z_dev, Z, (size_t)M * D * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    for (int d = 0; d < D; ++d) {
        double s = 0.0;
        for (int m = 0; m < M; ++m) s += Z[(size_t)m * D + d];
        ctx->center[d] = s / M;
    }
    SGP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->M = M; ctx->have_Z = true; ctx->have_kuu = false; ctx->have_stats = false;
    size_t need = (size_t)6 * M * M + 4 * (size_t)M + 64 + 2;      // (the last two M x M blocks: scratch of a K_uu job that shares the posterior's launch)
    return sgp_ensure(ctx, &ctx->dense_dev, &ctx->dense_cap, need);
}

static int alloc_data(sgp_ctx* ctx, int64_t N) {
    int64_t cap = ((N + 31) / 32) * 32;
    if (ctx->own_data && ctx->Ncap >= cap && ctx->X_dev && ctx->Dcap >= ctx->D) return SGP_OK;
    if (ctx->own_data) { cudaFree(ctx->X_dev); cudaFree(ctx->y_dev); cudaFree(ctx->yv_dev); cudaFree(ctx->w_dev); }
    ctx->X_dev = ctx->y_dev = ctx->yv_dev = ctx->w_dev = nullptr; ctx->own_data = true; ctx->Ncap = 0;
    SGP_CUDA(ctx, cudaMalloc((void**)&ctx->X_dev, (size_t)cap * ctx->D * sizeof(double)));
    SGP_CUDA(ctx, cudaMalloc((void**)&ctx->y_dev, (size_t)cap * sizeof(double)));
    SGP_CUDA(ctx, cudaMalloc((void**)&ctx->yv_dev, (size_t)cap * sizeof(double)));
    SGP_CUDA(ctx, cudaMalloc((void**)&ctx->w_dev, (size_t)cap * sizeof(double)));
    // zeroed ONCE: the sweep stages whole chunks of 32 points, and rows [N, cap) only have to hold finite values (they generate exact zeros)
    SGP_CUDA(ctx, cudaMemsetAsync(ctx->X_dev, 0, (size_t)cap * ctx->D * sizeof(double), ctx->stream));
    SGP_CUDA(ctx, cudaMemsetAsync(ctx->y_dev, 0, (size_t)cap * sizeof(double), ctx->stream));
    SGP_CUDA(ctx, cudaMemsetAsync(ctx->yv_dev, 0, (size_t)cap * sizeof(double), ctx->stream));
    SGP_CUDA(ctx, cudaMemsetAsync(ctx->w_dev, 0, (size_t)cap * sizeof(double), ctx->stream));
    ctx->Ncap = cap; ctx->Dcap = ctx->D;
    return SGP_OK;
}

static int upload_data(sgp_ctx* ctx, int64_t N, const double* X, const double* ybar, const double* yvar, const double* wts, bool beside = false);

int sgp_set_data(sgp_ctx* ctx, int64_t N, const double* X, const double* ybar, const double* yvar, const double* wts) {
    int rc = upload_data(ctx, N, X, ybar, yvar, wts); if (rc) return rc;
    SGP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));       // the host buffers may be reused as soon as this returns
    return SGP_OK;
}

// H2D copies enqueued on the ctx stream, no host synchronisation.
// beside: the copies go to the copy stream (ordered after everything already on the main stream) and end with the ready word; the main stream
// is NOT ordered after them -- the generate-once sweep launched next waits for the ready word on the device (its launch latency and set-up run
// under the copies), every other consumer calls sgp_join_upload first.
static int upload_data(sgp_ctx* ctx, int64_t N, const double* X, const double* ybar, const double* yvar, const double* wts, bool beside) {
    if (check(ctx)) return SGP_ERR_ARG;
    SGP_RANGE("sgp_upload");
    if (!ctx->have_kernel) SGP_FAIL(ctx, SGP_ERR_ARG, "set_data: set_kernel first (D is taken from it)");
    if (N < 0 || (N > 0 && !X)) SGP_FAIL(ctx, SGP_ERR_ARG, "set_data: X required");
    SGP_CUDA(ctx, cudaSetDevice(ctx->dev));
    if (!ctx->own_data) { ctx->X_dev = ctx->y_dev = ctx->yv_dev = ctx->w_dev = nullptr; ctx->Ncap = 0; }
    int rc = alloc_data(ctx, N > 0 ? N : 32); if (rc) return rc;
    const int D = ctx->D;
    const size_t cap = (size_t)ctx->Ncap;
    // the sweep stages whole chunks of 32 points: rows [N, cap) hold zeros from the allocation or finite values of an earlier, longer data
    // set -- either way they generate exact zeros (no per-upload memset)
    const size_t n = (size_t)N;
    cudaStream_t st = ctx->stream;
    if (beside && n) {
        if (!ctx->ready_dev) {
            SGP_CUDA(ctx, cudaMalloc((void**)&ctx->ready_dev, 64));
            SGP_CUDA(ctx, cudaMemsetAsync(ctx->ready_dev, 0, 64, ctx->stream));
            SGP_CUDA(ctx, cudaHostAlloc((void**)&ctx->ready_host, 8 * sizeof(unsigned), cudaHostAllocDefault));
        }
        SGP_CUDA(ctx, cudaEventRecord(ctx->ev_copy, ctx->stream));           // the previous sweep may still be reading the buffers
        SGP_CUDA(ctx, cudaStreamWaitEvent(ctx->stream2, ctx->ev_copy, 0));
        st = ctx->stream2;
    }
    if (n) SGP_CUDA(ctx, cudaMemcpyAsync(ctx->X_dev, X, n * D * sizeof(double), cudaMemcpyHostToDevice, st));
    if (ybar && n) SGP_CUDA(ctx, cudaMemcpyAsync(ctx->y_dev, ybar, n * sizeof(double), cudaMemcpyHostToDevice, st));
    else SGP_CUDA(ctx, cudaMemsetAsync(ctx->y_dev, 0, cap * sizeof(double), st));
    if (yvar && n) SGP_CUDA(ctx, cudaMemcpyAsync(ctx->yv_dev, yvar, n * sizeof(double), cudaMemcpyHostToDevice, st));
    if (wts && n) SGP_CUDA(ctx, cudaMemcpyAsync(ctx->w_dev, wts, n * sizeof(double), cudaMemcpyHostToDevice, st));
    if (st != ctx->stream) {
        // the ready word follows the data in stream order (one copy engine executes a stream's copies one after the other)
        const unsigned e = ++ctx->ready_epoch;
        ctx->ready_host[e & 7u] = e;
        SGP_CUDA(ctx, cudaMemcpyAsync(ctx->ready_dev, ctx->ready_host + (e & 7u), sizeof(unsigned), cudaMemcpyHostToDevice, st));
        ctx->upload_pending = true;
    }
    ctx->N = N; ctx->have_yv = yvar != nullptr; ctx->have_w = wts != nullptr; ctx->have_stats = false; ctx->have_data = true;
    return SGP_OK;
}

int sgp_set_targets(sgp_ctx* ctx, const double* ybar, const double* yvar) {
    if (check(ctx)) return SGP_ERR_ARG;
    if (!ctx->own_data || ctx->N <= 0) SGP_FAIL(ctx, SGP_ERR_ARG, "set_targets: needs data set with sgp_set_data");
    SGP_CUDA(ctx, cudaSetDevice(ctx->dev));
    const size_t N = (size_t)ctx->N;
    if (ybar) SGP_CUDA(ctx, cudaMemcpyAsync(ctx->y_dev, ybar, N * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    else SGP_CUDA(ctx, cudaMemsetAsync(ctx->y_dev, 0, N * sizeof(double), ctx->stream));
    if (yvar) SGP_CUDA(ctx, cudaMemcpyAsync(ctx->yv_dev, yvar, N * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    else SGP_CUDA(ctx, cudaMemsetAsync(ctx->yv_dev, 0, N * sizeof(double), ctx->stream));
    SGP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->have_yv = yvar != nullptr; ctx->have_stats = false;
    return SGP_OK;
}

int sgp_set_data_dev(sgp_ctx* ctx, int64_t N, const double* X_dev, const double* ybar_dev, const double* yvar_dev, const double* wts_dev) {
    if (check(ctx)) return SGP_ERR_ARG;
    if (!ctx->have_kernel) SGP_FAIL(ctx, SGP_ERR_ARG, "set_data_dev: set_kernel first");
    if (N <= 0 || !X_dev || !ybar_dev) SGP_FAIL(ctx, SGP_ERR_ARG, "set_data_dev: X_dev and ybar_dev required");
    if (N % 32) SGP_FAIL(ctx, SGP_ERR_ARG, "set_data_dev: N must be a multiple of 32 (buffers are used in place, unpadded)");
    SGP_CUDA(ctx, cudaSetDevice(ctx->dev));
    if (ctx->own_data) { cudaFree(ctx->X_dev); cudaFree(ctx->y_dev); cudaFree(ctx->yv_dev); cudaFree(ctx->w_dev); ctx->own_data = false; }
    ctx->X_dev = const_cast<double*>(X_dev); ctx->y_dev = const_cast<double*>(ybar_dev);
    ctx->yv_dev = const_cast<double*>(yvar_dev); ctx->w_dev = const_cast<double*>(wts_dev);
    ctx->N = N; ctx->Ncap = N; ctx->have_yv = yvar_dev != nullptr; ctx->have_w = wts_dev != nullptr; ctx->have_stats = false; ctx->have_data = true;
    return SGP_OK;
}

// packed[tri_col(j) + i - j] = A[i + j M], i >= j: the lower triangle column by column (LAPACK 'L' packed storage)
__global__ void pack_lower_kernel(const double* __restrict__ A, double* __restrict__ packed, int M) {
    const int j = blockIdx.x;
    const long long off = (long long)j * M - (long long)j * (j - 1) / 2 - j;
    for (int i = j + threadIdx.x; i < M; i += blockDim.x) packed[off + i] = A[(size_t)i + (size_t)j * M];
}

// packed: 0 = Psi2 as the full square; 1 = only the packed lower triangle of Psi2 into psi2 (M (M + 1) / 2 doubles); 2 = ALL statistics packed into psi2 --
// [lower triangle | Psi1 (M * D_out) | Psi0, sum_y2, sum_w, n], the layout of the multi-GPU exchange -- with ONE device-to-host copy
static int fetch_stats(sgp_ctx* ctx, double* psi0, double* psi1, double* psi2, double* sum_y2, int packed = 0) {
    SGP_RANGE("sgp_fetch");
    const size_t M = (size_t)ctx->M, Do = (size_t)ctx->Dout;
    double* s2 = ctx->stats_dev; double* s1 = s2 + M * M; double* sc = s1 + M * Do;
    const size_t small = M * Do + 4;
    if (psi2 && packed) {
        const size_t tri = M * (M + 1) / 2;
        const double* src = ctx->packed_src;
        if (!src) {      // the sweep did not leave a packed copy (first fused kernel, NCCL path): pack the resident statistics
            int rc = sgp_ensure(ctx, &ctx->packed_dev, &ctx->packed_cap, tri + small + 2); if (rc) return rc;
            pack_lower_kernel<<<(unsigned)M, 128, 0, ctx->stream>>>(s2, ctx->packed_dev, (int)M);
            SGP_CUDA(ctx, cudaGetLastError());
            if (packed == 2) SGP_CUDA(ctx, cudaMemcpyAsync(ctx->packed_dev + tri, s1, small * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
            src = ctx->packed_dev;
        }
        SGP_CUDA(ctx, cudaMemcpyAsync(psi2, src, (tri + (packed == 2 ? small : 0)) * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        if (packed == 2) {      // everything came with that one copy: the scalars and Psi1 are read from the tail of the caller's buffer
            SGP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            const double* tail = psi2 + tri;
            if (psi1) memcpy(psi1, tail, M * Do * sizeof(double));
            if (psi0) *psi0 = tail[M * Do];
            if (sum_y2) *sum_y2 = tail[M * Do + 1];
            return SGP_OK;
        }
    } else if (psi2) SGP_CUDA(ctx, cudaMemcpyAsync(psi2, s2, M * M * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    // Psi1 and the scalars are adjacent on the device: ONE copy into a pinned staging buffer, split on the host
    if (ctx->fetch_cap < small) {
        if (ctx->fetch_host) cudaFreeHost(ctx->fetch_host);
        ctx->fetch_host = nullptr; ctx->fetch_cap = 0;
        SGP_CUDA(ctx, cudaHostAlloc((void**)&ctx->fetch_host, (small + 1024) * sizeof(double), cudaHostAllocDefault));
        ctx->fetch_cap = small + 1024;
    }
    double* stage = ctx->fetch_host;
    if (psi1) SGP_CUDA(ctx, cudaMemcpyAsync(stage, s1, small * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    else SGP_CUDA(ctx, cudaMemcpyAsync(stage + M * Do, sc, 4 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    SGP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (psi1) memcpy(psi1, stage, M * Do * sizeof(double));
    if (psi0) *psi0 = stage[M * Do];
    if (sum_y2) *sum_y2 = stage[M * Do + 1];
    return SGP_OK;
}

}  // extern "C"
// orders the main stream after an upload that sgp_sweep_psi_host put on the copy stream (no-op otherwise)
int sgp_join_upload(sgp_ctx* ctx) {
    if (!ctx->upload_pending) return SGP_OK;
    ctx->upload_pending = false;
    SGP_CUDA(ctx, cudaEventRecord(ctx->ev_copy, ctx->stream2));
    SGP_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_copy, 0));
    return SGP_OK;
}
int sgp_sweep_resident(sgp_ctx* ctx, bool time_main) {
    if (!ctx->have_data) SGP_FAIL(ctx, SGP_ERR_ARG, "sweep: no data (sgp_set_data)");
    SGP_RANGE("sgp_sweep");
    ctx->stats_of_data = false;
    ctx->packed_src = nullptr;
    if (ctx->N == 0) {
        { int rcj = sgp_join_upload(ctx); if (rcj) return rcj; }
        // an empty data set (e.g. a rank whose shard is empty): all statistics are zero; under sharding the rank still takes part in the sum
        if (!ctx->have_kernel || !ctx->have_Z) SGP_FAIL(ctx, SGP_ERR_ARG, "sweep: set_kernel and set_inducing first");
        const size_t cnt = (size_t)ctx->M * ctx->M + (size_t)ctx->M + 4;
        int rc0 = sgp_ensure_stats(ctx, cnt + 4); if (rc0) return rc0;
        SGP_CUDA(ctx, cudaMemsetAsync(ctx->stats_dev, 0, cnt * sizeof(double), ctx->stream));
        ctx->Dout = 1; ctx->have_stats = true; ctx->last_launches = 0; ctx->last_sweep_exchanged = false;
        if (ctx->comm) { rc0 = sgp_comm_allreduce_stats(ctx, ctx->M, 1); if (rc0) return rc0; ctx->last_launches = 1; }
        ctx->stats_of_data = true;
        return SGP_OK;
    }
    ctx->want_exchange = true;
    int rc = sgp_sweep_launch(ctx, ctx->X_dev, ctx->y_dev, ctx->have_yv ? ctx->yv_dev : nullptr, ctx->have_w ? ctx->w_dev : nullptr, ctx->N,
                              ctx->Ncap, time_main);
    ctx->want_exchange = false;
    if (rc) return rc;
    if (ctx->comm && !ctx->last_sweep_exchanged) {       // (the generate-once kernel exchanges over peer memory inside its own launch)
        rc = sgp_comm_allreduce_stats(ctx, ctx->M, ctx->Dout); if (rc) return rc;
        ctx->last_launches += 1;
        ctx->packed_src = nullptr;   // (a packed copy written before the sum is stale)
    }
    ctx->stats_of_data = true;       // the resident statistics are those of the resident data (summed over the ranks)
    return SGP_OK;
}
extern "C" {

int sgp_sweep_psi(sgp_ctx* ctx, double* psi0, double* psi1, double* psi2, double* sum_y2) {
    if (check(ctx)) return SGP_ERR_ARG;
    SGP_CUDA(ctx, cudaSetDevice(ctx->dev));
    int rc = sgp_sweep_resident(ctx, false); if (rc) return rc;
    return fetch_stats(ctx, psi0, psi1, psi2, sum_y2);
}

static int sweep_host(sgp_ctx* ctx, int64_t N, const double* X, const double* ybar, const double* yvar, const double* wts, double* psi0,
                      double* psi1, double* psi2, double* sum_y2, bool packed) {
    const char* ov = std::getenv("SGP_HOST_OVERLAP");
    const bool beside = !(ov && ov[0] == '0');
    int rc = upload_data(ctx, N, X, ybar, yvar, wts, beside); if (rc) return rc;
    ctx->want_packed = packed && psi2 != nullptr;
    rc = sgp_sweep_resident(ctx, false);
    ctx->want_packed = false;
    const int rcj = sgp_join_upload(ctx);      // (a path that never launched a consumer of the ready word)
    if (rc) { cudaStreamSynchronize(ctx->stream2); return rc; }      // a failed sweep: nothing of this call stays in flight on the copy stream
    if (rcj) return rcj;
    return fetch_stats(ctx, psi0, psi1, psi2, sum_y2, packed ? 2 : 0);       // the one host synchronisation of the step
}

int sgp_sweep_psi_host(sgp_ctx* ctx, int64_t N, const double* X, const double* ybar, const double* yvar, const double* wts, double* psi0,
                       double* psi1, double* psi2, double* sum_y2) {
    return sweep_host(ctx, N, X, ybar, yvar, wts, psi0, psi1, psi2, sum_y2, false);
}

int sgp_sweep_psi_host_packed(sgp_ctx* ctx, int64_t N, const double* X, const double* ybar, const double* yvar, const double* wts, double* psi0,
                              double* psi1, double* stats_packed, double* sum_y2) {
    return sweep_host(ctx, N, X, ybar, yvar, wts, psi0, psi1, stats_packed, sum_y2, true);
}

int sgp_fetch_psi2_packed(sgp_ctx* ctx, double* psi2_packed) {
    if (check(ctx) || !psi2_packed) return SGP_ERR_ARG;
    if (!ctx->have_stats) SGP_FAIL(ctx, SGP_ERR_ARG, "fetch_psi2_packed: no statistics (sweep first)");
    SGP_CUDA(ctx, cudaSetDevice(ctx->dev));
    return fetch_stats(ctx, nullptr, nullptr, psi2_packed, nullptr, 1);
}

int sgp_sweep_psi_uncertain(sgp_ctx* ctx, int method, int p, int64_t N, const double* mean, const double* cov, int D_out, const double* R,
                            double* psi0, double* psi1, double* psi2, double* psi1_n) {
    if (check(ctx)) return SGP_ERR_ARG;
    SGP_CUDA(ctx, cudaSetDevice(ctx->dev));
    return sgp_uncertain_sweep(ctx, method, p, N, mean, cov, D_out, R, psi0, psi1, psi2, psi1_n);
}

int sgp_sweep_timed(sgp_ctx* ctx, int reps, float* ms_per_sweep, float* ms_main_kernel) {
    if (check(ctx)) return SGP_ERR_ARG;
    if (reps < 1) SGP_FAIL(ctx, SGP_ERR_ARG, "sweep_timed: reps >= 1");
    SGP_CUDA(ctx, cudaSetDevice(ctx->dev));
    SGP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    double main_sum = 0.0;
    SGP_CUDA(ctx, cudaEventRecord(ctx->ev[0], ctx->stream));
    for (int r = 0; r < reps; ++r) {
        int rc = sgp_sweep_resident(ctx, true); if (rc) return rc;
        // main-kernel time needs its own pair of events per repetition; read it back after the sync below for r == reps-1
        if (r + 1 < reps) {
            SGP_CUDA(ctx, cudaEventSynchronize(ctx->ev[3]));
            float ms = 0.f; SGP_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->ev[2], ctx->ev[3])); main_sum += ms;
        }
    }
    SGP_CUDA(ctx, cudaEventRecord(ctx->ev[1], ctx->stream));
    SGP_CUDA(ctx, cudaEventSynchronize(ctx->ev[1]));
    float ms = 0.f; SGP_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->ev[2], ctx->ev[3])); main_sum += ms;
    float tot = 0.f; SGP_CUDA(ctx, cudaEventElapsedTime(&tot, ctx->ev[0], ctx->ev[1]));
    if (ms_per_sweep) *ms_per_sweep = tot / reps;
    if (ms_main_kernel) *ms_main_kernel = (float)(main_sum / reps);
    ctx->last_main_ms = (float)(main_sum / reps);
    return SGP_OK;
}

int sgp_sweep_timed_flushed(sgp_ctx* ctx, int reps, int flush_mb, float* ms_per_sweep, float* ms_main_kernel) {
    if (check(ctx)) return SGP_ERR_ARG;
    if (reps < 1 || reps > 4096 || flush_mb < 0) SGP_FAIL(ctx, SGP_ERR_ARG, "sweep_timed_flushed: 1 <= reps <= 4096, flush_mb >= 0");
    SGP_CUDA(ctx, cudaSetDevice(ctx->dev));
    const size_t fbytes = (size_t)flush_mb << 20;
    if (fbytes > ctx->flush_cap) {
        if (ctx->flush_dev) SGP_CUDA(ctx, cudaFree(ctx->flush_dev));
        ctx->flush_dev = nullptr; ctx->flush_cap = 0;
        SGP_CUDA(ctx, cudaMalloc(&ctx->flush_dev, fbytes));
        ctx->flush_cap = fbytes;
    }
    // everything is enqueued without a host synchronisation: [flush] [e0] sweep (+ exchange) [e1] per repetition, so that the ranks of a
    // multi-GPU job stay in lock step through the exchange itself instead of accumulating host-side launch skew
    // ms_main_kernel == NULL: no inner event pair around the main kernel -- the timed interval then holds the sweep and nothing else (each event
    // record costs about a microsecond of stream time; the headline measurement runs without them)
    const bool inner = ms_main_kernel != nullptr;
    std::vector<cudaEvent_t> ev(4 * (size_t)reps, nullptr);
    for (auto& e : ev) SGP_CUDA(ctx, cudaEventCreate(&e));
    cudaEvent_t keep2 = ctx->ev[2], keep3 = ctx->ev[3];
    int rc = SGP_OK;
    for (int r = 0; r < reps && rc == SGP_OK; ++r) {
        if (fbytes && cudaMemsetAsync(ctx->flush_dev, r & 0xff, fbytes, ctx->stream) != cudaSuccess) { rc = SGP_ERR_CUDA; break; }
        // the flush takes a few microseconds more on one GPU than on another: align the ranks again BEFORE the timed interval starts (a flag
        // barrier over peer memory on the stream), otherwise that skew is waited out inside the timed sweep's exchange
        if (ctx->comm && sgp_comm_barrier(ctx) != SGP_OK) { rc = SGP_ERR_CUDA; break; }
        if (cudaEventRecord(ev[4 * r], ctx->stream) != cudaSuccess) { rc = SGP_ERR_CUDA; break; }
        ctx->ev[2] = ev[4 * r + 2]; ctx->ev[3] = ev[4 * r + 3];       // the launcher brackets the main kernel with ev[2] / ev[3]
        rc = sgp_sweep_resident(ctx, inner);
        if (rc == SGP_OK && cudaEventRecord(ev[4 * r + 1], ctx->stream) != cudaSuccess) rc = SGP_ERR_CUDA;
    }
    ctx->ev[2] = keep2; ctx->ev[3] = keep3;
    double tot = 0.0, main_sum = 0.0;
    if (rc == SGP_OK && cudaStreamSynchronize(ctx->stream) != cudaSuccess) rc = SGP_ERR_CUDA;
    for (int r = 0; r < reps && rc == SGP_OK; ++r) {
        float a = 0.f, b = 0.f;
        if (cudaEventElapsedTime(&a, ev[4 * r], ev[4 * r + 1]) != cudaSuccess || (inner && cudaEventElapsedTime(&b, ev[4 * r + 2], ev[4 * r + 3]) != cudaSuccess)) rc = SGP_ERR_CUDA;
        tot += a; main_sum += b;
    }
    for (auto& e : ev) if (e) cudaEventDestroy(e);
    if (rc != SGP_OK) { if (ctx->err.empty()) ctx->err = "sweep_timed_flushed: CUDA failure"; return rc; }
    if (ms_per_sweep) *ms_per_sweep = (float)(tot / reps);
    if (ms_main_kernel) { *ms_main_kernel = (float)(main_sum / reps); ctx->last_main_ms = (float)(main_sum / reps); }
    return SGP_OK;
}

int sgp_last_sweep_info(sgp_ctx* ctx, int* n_launches, int* grid, int* block, int* smem_bytes) {
    if (check(ctx)) return SGP_ERR_ARG;
    if (n_launches) *n_launches = ctx->last_launches;
    if (grid) *grid = ctx->last_grid;
    if (block) *block = ctx->last_block;
    if (smem_bytes) *smem_bytes = ctx->last_smem;
    return SGP_OK;
}

int sgp_sweep_debug_clocks(sgp_ctx* ctx, int64_t* out, int cap, int* nrec) {
    if (check(ctx)) return SGP_ERR_ARG;
    SGP_CUDA(ctx, cudaSetDevice(ctx->dev));
    constexpr int kMaxSlots = 8192;
    if (nrec) *nrec = 0;
    if (!ctx->sweep_dbg_dev) {
        SGP_CUDA(ctx, cudaMalloc(&ctx->sweep_dbg_dev, (size_t)kMaxSlots * 4 * sizeof(long long)));
        SGP_CUDA(ctx, cudaMemsetAsync(ctx->sweep_dbg_dev, 0xff, (size_t)kMaxSlots * 4 * sizeof(long long), ctx->stream));
        return SGP_OK;
    }
    int n = ctx->sweep_dbg_slots < cap ? ctx->sweep_dbg_slots : cap;
    if (n > 0 && out) {
        SGP_CUDA(ctx, cudaMemcpyAsync(out, ctx->sweep_dbg_dev, (size_t)n * 4 * sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
        SGP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    if (nrec) *nrec = n;
    return SGP_OK;
}

int sgp_stats_dev(sgp_ctx* ctx, double** psi2_dev, double** psi1_dev, double** scal_dev) {
    if (check(ctx)) return SGP_ERR_ARG;
    if (!ctx->have_stats) SGP_FAIL(ctx, SGP_ERR_ARG, "stats_dev: no sweep yet");
    const size_t M = (size_t)ctx->M;
    if (psi2_dev) *psi2_dev = ctx->stats_dev;
    if (psi1_dev) *psi1_dev = ctx->stats_dev + M * M;
    if (scal_dev) *scal_dev = ctx->stats_dev + M * M + M * ctx->Dout;
    return SGP_OK;
}

// ---- M x M factorisations ---------------------------------------------------------------------------------------
}  // extern "C"
// Enqueues the K_uu job without synchronising: the caller checks the pivot record (sgp_dense_info, or its own read of ctx->info_dev) when it
// next synchronises, and clears ctx->have_kuu if that fails.
static int kuu_job_prepare(sgp_ctx* ctx, double jitter, double* scratch /* 2 M x M */, SgpDenseJob& j) {
    const int M = ctx->M;
    if (ctx->KuuL_M != M) {
        if (ctx->KuuL_dev) SGP_CUDA(ctx, cudaFree(ctx->KuuL_dev));
        if (ctx->Kinv_dev) SGP_CUDA(ctx, cudaFree(ctx->Kinv_dev));
        ctx->KuuL_dev = nullptr; ctx->Kinv_dev = nullptr;
        SGP_CUDA(ctx, cudaMalloc((void**)&ctx->KuuL_dev, (size_t)M * M * sizeof(double)));
        SGP_CUDA(ctx, cudaMalloc((void**)&ctx->Kinv_dev, (size_t)M * M * sizeof(double)));
        ctx->KuuL_M = M;
    }
    ctx->have_kuu = false;
    // ONE cooperative launch: K_uu = kernelmatrix(Xu) + jitter I -> L (and the inverses of its diagonal blocks, which sgp_kuu_solve uses)
    // -> X = L^-1 -> K_uu^-1 = X' X, which the :w rule / energy / theta step contract with Psi2
    const size_t nd = (size_t)((M + 63) / 64) * 64 * 64, MM = (size_t)M * M;
    int rc = sgp_ensure_zero(ctx, &ctx->kuu_dinv_dev, &ctx->kuu_dinv_cap, nd); if (rc) return rc;
    j = SgpDenseJob{};
    j.M = M; j.build = 2; j.jitter = jitter; j.A = ctx->KuuL_dev; j.Dinv = ctx->kuu_dinv_dev; j.X = scratch; j.Tmp = scratch + MM; j.S = ctx->Kinv_dev;
    return SGP_OK;
}
int sgp_kuu_factor_enqueue(sgp_ctx* ctx, double jitter) {
    SgpDenseJob j;
    int rc = kuu_job_prepare(ctx, jitter, ctx->dense_dev + 64, j); if (rc) return rc;
    rc = sgp_dense_job(ctx, j); if (rc) return rc;
    ctx->have_kuu = true; ctx->kuu_jitter = jitter; ctx->kuu_jitter_known = true;
    return SGP_OK;
}
extern "C" {

int sgp_kuu_factor(sgp_ctx* ctx, double jitter, double* L) {
    if (check(ctx)) return SGP_ERR_ARG;
    if (!ctx->have_kernel || !ctx->have_Z) SGP_FAIL(ctx, SGP_ERR_ARG, "kuu_factor: set_kernel and set_inducing first");
    SGP_RANGE("sgp_kuu_factor");
    SGP_CUDA(ctx, cudaSetDevice(ctx->dev));
    int rc = sgp_kuu_factor_enqueue(ctx, jitter); if (rc) return rc;
    if (ctx->dense_timing) SGP_CUDA(ctx, cudaEventRecord(ctx->ev[1], ctx->stream));
    if (L) SGP_CUDA(ctx, cudaMemcpyAsync(L, ctx->KuuL_dev, (size_t)ctx->M * ctx->M * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    rc = sgp_dense_info(ctx, "kuu_factor");
    if (rc) { ctx->have_kuu = false; return rc; }
    return SGP_OK;
}

int sgp_kuu_solve(sgp_ctx* ctx, int nrhs, double* B) {
    if (check(ctx)) return SGP_ERR_ARG;
    if (!ctx->have_kuu) SGP_FAIL(ctx, SGP_ERR_ARG, "kuu_solve: sgp_kuu_factor first");
    if (nrhs < 1 || !B) SGP_FAIL(ctx, SGP_ERR_ARG, "kuu_solve: nrhs >= 1 and B required");
    SGP_CUDA(ctx, cudaSetDevice(ctx->dev));
    const int M = ctx->M;
    double* tmp = nullptr;
    size_t bytes = (size_t)M * nrhs * sizeof(double);
    SGP_CUDA(ctx, cudaMalloc((void**)&tmp, bytes + (size_t)64 * nrhs * sizeof(double)));
    double* panel = tmp + (size_t)M * nrhs;
    int rc = SGP_OK;
    if (cudaMemcpyAsync(tmp, B, bytes, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess) rc = SGP_ERR_CUDA;
    if (!rc) rc = sgp_trsm_lower_dinv(ctx, ctx->KuuL_dev, ctx->kuu_dinv_dev, tmp, panel, M, nrhs, false);
    if (!rc) rc = sgp_trsm_lower_dinv(ctx, ctx->KuuL_dev, ctx->kuu_dinv_dev, tmp, panel, M, nrhs, true);
    if (!rc && cudaMemcpyAsync(B, tmp, bytes, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess) rc = SGP_ERR_CUDA;
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess && !rc) rc = SGP_ERR_CUDA;
    cudaFree(tmp);
    if (rc == SGP_ERR_CUDA && ctx->err.empty()) ctx->err = "kuu_solve: CUDA failure";
    return rc;
}

// resident prior / posterior of v: post_dev = [Lambda_prior (MM) | xi_prior (M) | mu (M) | Sigma (MM) | Uv (MM, upper)]
static int ensure_post(sgp_ctx* ctx) {
    const size_t M = (size_t)ctx->M, need = 3 * M * M + 2 * M;
    if (ctx->post_M != ctx->M) { ctx->have_prior = ctx->have_post = false; ctx->post_M = ctx->M; }
    return sgp_ensure(ctx, &ctx->post_dev, &ctx->post_cap, need);
}
static double* post_LamP(sgp_ctx* ctx) { return ctx->post_dev; }
static double* post_xiP(sgp_ctx* ctx) { return ctx->post_dev + (size_t)ctx->M * ctx->M; }
static double* post_mu(sgp_ctx* ctx) { return post_xiP(ctx) + ctx->M; }
static double* post_Sig(sgp_ctx* ctx) { return post_mu(ctx) + ctx->M; }
static double* post_Uv(sgp_ctx* ctx) { return post_Sig(ctx) + (size_t)ctx->M * ctx->M; }

__global__ void isotropic_kernel(double* __restrict__ A, int M, double diag) {
    size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (e < (size_t)M * M) A[e] = (e % M == e / M) ? diag : 0.0;
}

// The N-th prod on the resident prior: Lambda = Lambda_p + w Psi2, xi = xi_p + w Psi1 -> Sigma, mu (and, when the host wants it, Uv).
// carry: the posterior's natural parameters become the resident prior (the streaming schedule of regression_kin40k.ipynb:200-213).
// ONE cooperative launch: [J Lambda J -> L -> X = L^-1 -> Sigma = J X'X J -> mu -> Uv = G (J X J)] (dense_coop.cu).  The older two-launch
// form [Lambda -> ... -> mu] + [Sigma + mu mu' -> its Cholesky factor -> Uv] remains for M > SGP_FLIP_UV_MAX_M and as a cross-check
// (SGP_DENSE_UV2=1); there the device-to-host copies of Sigma / mu run on the second stream while the second launch factorises.
static int posterior_core(sgp_ctx* ctx, double w, bool carry, bool want_uv, double* mu_v, double* Sigma_v, double* Uv_host) {
    const int M = ctx->M; const size_t MM = (size_t)M * M;
    double* d = ctx->dense_dev + 64;
    double *Lam = d, *X = d + MM, *T = d + 3 * MM, *xi = d + 4 * MM;
    double* psi2 = ctx->stats_dev; double* psi1 = psi2 + MM;
    int rc = sgp_ensure_zero(ctx, &ctx->dinv_dev, &ctx->dinv_cap, (size_t)((M + 63) / 64) * 64 * 64); if (rc) return rc;
    // Reversed-order factorisation (default): the inverse of that factor IS the Cholesky factor of Sigma, and Uv follows from it by the
    // closed-form rank-one update inside the same launch.  SGP_DENSE_UV2=1 (or M > SGP_FLIP_UV_MAX_M) keeps the second factorisation.
    static const bool two_jobs = std::getenv("SGP_DENSE_UV2") != nullptr && std::atoi(std::getenv("SGP_DENSE_UV2")) != 0;
    const bool flip = !two_jobs && M <= SGP_FLIP_UV_MAX_M;
    SgpDenseJob a;
    a.M = M; a.build = 1; a.S2 = psi2; a.s1 = psi1; a.P = post_LamP(ctx); a.xip = post_xiP(ctx); a.xi = xi; a.w = w; a.carry = carry ? 1 : 0;
    a.A = Lam; a.Dinv = ctx->dinv_dev; a.X = X; a.Tmp = T; a.S = post_Sig(ctx); a.mu = post_mu(ctx);
    if (flip) { a.flip = 1; a.S = d + 2 * MM; a.Sout = post_Sig(ctx); if (want_uv) a.Uv = post_Uv(ctx); }
    // K_uu stale (new theta or inducing points) and a jitter on record: its factorisation + inverse share this launch -- a job is a serial chain on
    // ONE CTA with the others mostly waiting, so the second job costs next to nothing (kin40k mini-batch schedule: prod, then the theta step).
    // Speculative: a non-positive pivot there is not this call's error (the next sgp_kuu_factor / theta step reports it).
    static const bool no_fuse = std::getenv("SGP_DENSE_FUSE_KUU") != nullptr && std::atoi(std::getenv("SGP_DENSE_FUSE_KUU")) == 0;
    const bool with_kuu = !no_fuse && !ctx->have_kuu && ctx->kuu_jitter_known && ctx->have_kernel && ctx->have_Z;
    SgpDenseJob kj;
    if (with_kuu) { rc = kuu_job_prepare(ctx, ctx->kuu_jitter, d + 4 * MM + ((4 * (size_t)M + 1) & ~(size_t)1), kj); if (rc) return rc; }
    rc = sgp_dense_job(ctx, a, with_kuu ? &kj : nullptr); if (rc) return rc;
    ctx->kuu_speculative = with_kuu;
    ctx->have_post = true; ctx->have_post_uv = flip && want_uv;
    if (flip) want_uv = false;                                  // (done)
    const bool early = (Sigma_v || mu_v) && want_uv;            // something to copy while the second factorisation runs
    if (early) {
        SGP_CUDA(ctx, cudaEventRecord(ctx->ev_copy, ctx->stream));
        SGP_CUDA(ctx, cudaStreamWaitEvent(ctx->stream2, ctx->ev_copy, 0));
        if (Sigma_v) SGP_CUDA(ctx, cudaMemcpyAsync(Sigma_v, post_Sig(ctx), MM * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream2));
        if (mu_v) SGP_CUDA(ctx, cudaMemcpyAsync(mu_v, post_mu(ctx), (size_t)M * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream2));
    }
    if (want_uv) {
        SgpDenseJob b;
        b.M = M; b.build = 3; b.Sig = post_Sig(ctx); b.mu_in = post_mu(ctx); b.A = X; b.Dinv = ctx->dinv_dev; b.Ut = post_Uv(ctx);
        b.reset_info = 0;                                       // keep a non-positive pivot of the first factorisation on record
        rc = sgp_dense_job(ctx, b); if (rc) return rc;
        ctx->have_post_uv = true;
    }
    if (ctx->dense_timing) SGP_CUDA(ctx, cudaEventRecord(ctx->ev[1], ctx->stream));
    if (Uv_host) SGP_CUDA(ctx, cudaMemcpyAsync(Uv_host, post_Uv(ctx), MM * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    if (!early) {
        if (Sigma_v) SGP_CUDA(ctx, cudaMemcpyAsync(Sigma_v, post_Sig(ctx), MM * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        if (mu_v) SGP_CUDA(ctx, cudaMemcpyAsync(mu_v, post_mu(ctx), (size_t)M * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    }
    rc = sgp_dense_info(ctx, "posterior_v");                   // the one host synchronisation of the call
    if (ctx->kuu_speculative) {
        ctx->kuu_speculative = false;
        if (rc == SGP_OK && ctx->info2_last == 0) ctx->have_kuu = true;      // (jitter unchanged: ctx->kuu_jitter)
    }
    if (early && cudaStreamSynchronize(ctx->stream2) != cudaSuccess && rc == SGP_OK) { ctx->err = "posterior_v: copy stream failure"; rc = SGP_ERR_CUDA; }
    if (rc) { ctx->have_post = ctx->have_post_uv = false; }
    return rc;
}

}  // extern "C"
const double* sgp_resident_mu(sgp_ctx* ctx) { return (ctx->have_post && ctx->post_M == ctx->M) ? post_mu(ctx) : nullptr; }
const double* sgp_resident_sigma(sgp_ctx* ctx) { return (ctx->have_post && ctx->post_M == ctx->M) ? post_Sig(ctx) : nullptr; }
const double* sgp_resident_uv(sgp_ctx* ctx) { return (ctx->have_post && ctx->have_post_uv && ctx->post_M == ctx->M) ? post_Uv(ctx) : nullptr; }
extern "C" {

int sgp_prior_set(sgp_ctx* ctx, const double* xi0, const double* Lambda0) {
    if (check(ctx)) return SGP_ERR_ARG;
    if (!ctx->have_Z || !xi0 || !Lambda0) SGP_FAIL(ctx, SGP_ERR_ARG, "prior_set: set_inducing first; xi0 and Lambda0 required");
    SGP_CUDA(ctx, cudaSetDevice(ctx->dev));
    int rc = ensure_post(ctx); if (rc) return rc;
    const size_t M = (size_t)ctx->M;
    SGP_CUDA(ctx, cudaMemcpyAsync(post_LamP(ctx), Lambda0, M * M * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    SGP_CUDA(ctx, cudaMemcpyAsync(post_xiP(ctx), xi0, M * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    SGP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->have_prior = true;
    return SGP_OK;
}

int sgp_prior_set_isotropic(sgp_ctx* ctx, double variance) {
    if (check(ctx)) return SGP_ERR_ARG;
    if (!ctx->have_Z || !(variance > 0.0)) SGP_FAIL(ctx, SGP_ERR_ARG, "prior_set_isotropic: set_inducing first; variance > 0");
    SGP_CUDA(ctx, cudaSetDevice(ctx->dev));
    int rc = ensure_post(ctx); if (rc) return rc;
    const int M = ctx->M;
    isotropic_kernel<<<nblocks((size_t)M * M), 256, 0, ctx->stream>>>(post_LamP(ctx), M, 1.0 / variance);
    SGP_CUDA(ctx, cudaMemsetAsync(post_xiP(ctx), 0, (size_t)M * sizeof(double), ctx->stream));
    SGP_CUDA(ctx, cudaGetLastError());
    ctx->have_prior = true;
    return SGP_OK;
}

int sgp_posterior_v_stream(sgp_ctx* ctx, double w, int carry, double* mu_v, double* Sigma_v, double* Uv) {
    if (check(ctx)) return SGP_ERR_ARG;
    if (!ctx->have_stats) SGP_FAIL(ctx, SGP_ERR_ARG, "posterior_v_stream: run a sweep first");
    if (ctx->Dout != 1) SGP_FAIL(ctx, SGP_ERR_UNSUPPORTED, "posterior_v_stream: scalar-output statistics only");
    SGP_RANGE("sgp_posterior_v_stream");
    SGP_CUDA(ctx, cudaSetDevice(ctx->dev));
    int rc = ensure_post(ctx); if (rc) return rc;
    if (!ctx->have_prior) SGP_FAIL(ctx, SGP_ERR_ARG, "posterior_v_stream: sgp_prior_set / sgp_prior_set_isotropic first");
    return posterior_core(ctx, w, carry != 0, Uv != nullptr, mu_v, Sigma_v, Uv);
}

int sgp_posterior_v(sgp_ctx* ctx, const double* xi0, const double* Lambda0, double w, double* mu_v, double* Sigma_v, double* Uv) {
    if (check(ctx)) return SGP_ERR_ARG;
    if (!ctx->have_stats) SGP_FAIL(ctx, SGP_ERR_ARG, "posterior_v: run a sweep first");
    if (ctx->Dout != 1) SGP_FAIL(ctx, SGP_ERR_UNSUPPORTED, "posterior_v: scalar-output statistics only (MultiSGP folds kron(W, Psi2) on the host)");
    if (!xi0 || !Lambda0) SGP_FAIL(ctx, SGP_ERR_ARG, "posterior_v: prior natural parameters required");
    SGP_RANGE("sgp_posterior_v");
    SGP_CUDA(ctx, cudaSetDevice(ctx->dev));
    int rc = ensure_post(ctx); if (rc) return rc;
    const size_t M = (size_t)ctx->M;
    // the prior's natural parameters: enqueued, no host synchronisation of their own (this call returns only after its final one)
    SGP_CUDA(ctx, cudaMemcpyAsync(post_LamP(ctx), Lambda0, M * M * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    SGP_CUDA(ctx, cudaMemcpyAsync(post_xiP(ctx), xi0, M * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    ctx->have_prior = true;
    return posterior_core(ctx, w, false, Uv != nullptr, mu_v, Sigma_v, Uv);
}

int sgp_w_terms(sgp_ctx* ctx, const double* mu_v, const double* Uv, double* sumI1, double* sumI2) {
    if (check(ctx)) return SGP_ERR_ARG;
    if (!ctx->have_stats) SGP_FAIL(ctx, SGP_ERR_ARG, "w_terms: run a sweep first");
    if (!ctx->have_kuu) SGP_FAIL(ctx, SGP_ERR_ARG, "w_terms: sgp_kuu_factor first");
    if (ctx->Dout != 1) SGP_FAIL(ctx, SGP_ERR_UNSUPPORTED, "w_terms: scalar-output statistics only");
    if ((mu_v == nullptr) != (Uv == nullptr)) SGP_FAIL(ctx, SGP_ERR_ARG, "w_terms: pass both mu_v and Uv, or neither (resident posterior)");
    if (!mu_v && !sgp_resident_mu(ctx)) SGP_FAIL(ctx, SGP_ERR_ARG, "w_terms: no resident posterior (sgp_posterior_v first)");
    SGP_RANGE("sgp_w_terms");
    SGP_CUDA(ctx, cudaSetDevice(ctx->dev));
    const int M = ctx->M; const size_t MM = (size_t)M * M;
    double* d = ctx->dense_dev + 64;
    double *U = d + MM, *R = d + 2 * MM, *mu = d + 4 * MM, *res = mu + M;
    double* psi2 = ctx->stats_dev; double* psi1 = psi2 + MM; double* scal = psi1 + M;
    int rc = SGP_OK;
    if (mu_v) {
        // the host's (previous sweep's) posterior: R_v = Uv' Uv, then <K_uu^-1, Psi2>, <R_v, Psi2>, mu' Psi1 in one pass
        SGP_CUDA(ctx, cudaMemcpyAsync(U, Uv, MM * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        SGP_CUDA(ctx, cudaMemcpyAsync(mu, mu_v, M * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        rc = sgp_gemm(ctx, 1, 0, M, M, M, 1.0, U, M, U, M, 0.0, R, M, 0); if (rc) return rc;
        rc = sgp_wterms_reduce(ctx, ctx->Kinv_dev, psi2, R, nullptr, mu, psi1, M, res); if (rc) return rc;
    } else {
        // resident posterior: <R_v, Psi2> = <Sigma_v, Psi2> + mu_v' Psi2 mu_v -- no Cholesky factor of R_v needed
        rc = sgp_wterms_reduce(ctx, ctx->Kinv_dev, psi2, sgp_resident_sigma(ctx), sgp_resident_mu(ctx), sgp_resident_mu(ctx), psi1, M, res); if (rc) return rc;
    }
    if (ctx->dense_timing) SGP_CUDA(ctx, cudaEventRecord(ctx->ev[1], ctx->stream));
    double h[4], sc[4];
    SGP_CUDA(ctx, cudaMemcpyAsync(h, res, 4 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    SGP_CUDA(ctx, cudaMemcpyAsync(sc, scal, 4 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    SGP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (sumI1) *sumI1 = sc[0] - h[0];
    if (sumI2) *sumI2 = sc[1] - 2.0 * h[2] + h[1];
    return SGP_OK;
}

static int predict_impl(sgp_ctx* ctx, const char* what, int64_t Nt, const double* Xt, const double* mu_v, double w_bar, double* out, double* prob) {
    if (check(ctx)) return SGP_ERR_ARG;
    if (!ctx->have_kernel || !ctx->have_Z) SGP_FAIL(ctx, SGP_ERR_ARG, std::string(what) + ": set_kernel and set_inducing first");
    if (Nt < 1 || !Xt || !mu_v || !out) SGP_FAIL(ctx, SGP_ERR_ARG, std::string(what) + ": bad arguments");
    if (prob && !(w_bar > 0.0)) SGP_FAIL(ctx, SGP_ERR_ARG, std::string(what) + ": w_bar > 0 required");
    SGP_RANGE("sgp_predict");
    SGP_CUDA(ctx, cudaSetDevice(ctx->dev));
    const int M = ctx->M, D = ctx->D;
    // scratch from the context's arena: [Xt (Nt x D) | out (Nt) | prob (Nt) | mu (M) | 1 / ell (D)]
    const size_t need = (size_t)Nt * D + 2 * (size_t)Nt + (size_t)M + SGP_MAX_D + 8;
    int rc = sgp_ensure(ctx, &ctx->pred_dev, &ctx->pred_cap, need); if (rc) return rc;
    double* xt = ctx->pred_dev; double* o = xt + (size_t)Nt * D; double* pr = o + Nt; double* mu = pr + Nt;
    double inv[SGP_MAX_D];
    for (int d = 0; d < D; ++d) inv[d] = 1.0 / ctx->ell[d];
    SGP_CUDA(ctx, cudaMemcpyAsync(xt, Xt, (size_t)Nt * D * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    SGP_CUDA(ctx, cudaMemcpyAsync(mu, mu_v, (size_t)M * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    SGP_CUDA(ctx, cudaMemcpyAsync(mu + M, inv, D * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    const size_t sh = (size_t)256 * (D + 1) * sizeof(double);
    predict_kernel<<<nblocks((size_t)Nt, 128), 128, sh, ctx->stream>>>(xt, ctx->Z_dev, mu, o, Nt, M, D, ctx->kind, ctx->variance, mu + M, prob ? pr : nullptr,
                                                                       prob ? 1.0 / std::sqrt(1.0 + 1.0 / w_bar) : 0.0);
    SGP_CUDA(ctx, cudaGetLastError());
    SGP_CUDA(ctx, cudaMemcpyAsync(out, o, (size_t)Nt * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    if (prob) SGP_CUDA(ctx, cudaMemcpyAsync(prob, pr, (size_t)Nt * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    SGP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));       // (inv[] is a stack buffer)
    return SGP_OK;
}

int sgp_predict_mean(sgp_ctx* ctx, int64_t Nt, const double* Xt, const double* mu_v, double* out) {
    return predict_impl(ctx, "predict_mean", Nt, Xt, mu_v, 1.0, out, nullptr);
}

int sgp_predict_probit(sgp_ctx* ctx, int64_t Nt, const double* Xt, const double* mu_v, double w_bar, double* mean_f, double* var_f, double* prob_y) {
    if (!prob_y) { if (ctx) ctx->err = "predict_probit: prob_y required"; return SGP_ERR_ARG; }
    int rc = predict_impl(ctx, "predict_probit", Nt, Xt, mu_v, w_bar, mean_f, prob_y); if (rc) return rc;
    if (var_f) *var_f = 1.0 / w_bar;
    return SGP_OK;
}


// ---- measurement hooks of the M x M path (bench.py `dense` leg) ---------------------------------------------------------------------
// Device time (CUDA events on the ctx stream, kernels only: from the call's first enqueue to its last kernel) of `reps` calls of
//   what = 0: sgp_kuu_factor(jitter)   1: sgp_posterior_v_stream(w, carry = 0) with Uv   2: ... without Uv   3: sgp_w_terms (resident posterior)
// on the resident state (needs a sweep; 1-3 need a resident prior, 3 a posterior and K_uu).  Returns the mean per call.
int sgp_dense_timed(sgp_ctx* ctx, int what, double w, double jitter, int reps, float* ms_per_call) {
    if (check(ctx)) return SGP_ERR_ARG;
    if (reps < 1 || what < 0 || what > 3 || !ms_per_call) SGP_FAIL(ctx, SGP_ERR_ARG, "dense_timed: bad arguments");
    SGP_CUDA(ctx, cudaSetDevice(ctx->dev));
    double tot = 0.0;
    int rc = SGP_OK;
    for (int r = 0; r < reps && rc == SGP_OK; ++r) {
        SGP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        ctx->dense_timing = true;
        if (cudaEventRecord(ctx->ev[0], ctx->stream) != cudaSuccess) rc = SGP_ERR_CUDA;
        double a = 0.0, b = 0.0;
        if (rc == SGP_OK) {
            if (what == 0) rc = sgp_kuu_factor(ctx, jitter, nullptr);
            else if (what == 1 || what == 2) {
                if (!ctx->have_stats || !ctx->have_prior || ctx->Dout != 1) { ctx->err = "dense_timed: needs a sweep and a resident prior"; rc = SGP_ERR_ARG; }
                else { rc = ensure_post(ctx); if (rc == SGP_OK) rc = posterior_core(ctx, w, false, what == 1, nullptr, nullptr, nullptr); }
            } else rc = sgp_w_terms(ctx, nullptr, nullptr, &a, &b);
        }
        ctx->dense_timing = false;
        if (rc == SGP_OK) {
            float ms = 0.f;
            if (cudaEventSynchronize(ctx->ev[1]) != cudaSuccess || cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]) != cudaSuccess) rc = SGP_ERR_CUDA;
            tot += ms;
        }
    }
    if (rc == SGP_ERR_CUDA && ctx->err.empty()) ctx->err = "dense_timed: CUDA failure";
    if (rc) return rc;
    *ms_per_call = (float)(tot / reps);
    return SGP_OK;
}

}  // extern "C"

// FP64 tensor-pipe peak measured on THIS device: every warp keeps 16 independent DMMA.8x8x4 accumulator chains busy (registers only)
__global__ void __launch_bounds__(256) fp64_peak_kernel(double* __restrict__ out, int iters, double seed) {
    double c[16][2];
#pragma unroll
    for (int i = 0; i < 16; ++i) c[i][0] = c[i][1] = 0.0;
    const double a = seed * (threadIdx.x + 1) * 1e-3, b = seed * (threadIdx.x + 3) * 1e-3;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) dmma884(c[i][0], c[i][1], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += c[i][0] + c[i][1];
    if (s == 12345.678) out[0] = s;       // never true: keeps the chains alive
}

extern "C" int sgp_fp64_peak(sgp_ctx* ctx, int ms_target, double* dmma_tflops) {
    if (check(ctx)) return SGP_ERR_ARG;
    if (!dmma_tflops || ms_target < 1) SGP_FAIL(ctx, SGP_ERR_ARG, "fp64_peak: bad arguments");
    SGP_CUDA(ctx, cudaSetDevice(ctx->dev));
    const int grid = 4 * ctx->num_sms;
    int iters = 20000;
    double best = 0.0;
    for (int pass = 0; pass < 4; ++pass) {
        SGP_CUDA(ctx, cudaEventRecord(ctx->ev[0], ctx->stream));
        fp64_peak_kernel<<<grid, 256, 0, ctx->stream>>>(ctx->exptab_dev, iters, 1.0e-3);
        SGP_CUDA(ctx, cudaEventRecord(ctx->ev[1], ctx->stream));
        SGP_CUDA(ctx, cudaEventSynchronize(ctx->ev[1]));
        float ms = 0.f; SGP_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]));
        const double flops = (double)grid * 8.0 * iters * 16.0 * 512.0;          // warps x iterations x chains x (8 x 8 x 4 FMA = 512 FLOP)
        if (pass > 0) best = std::max(best, flops / (ms * 1e-3) * 1e-12);
        if (pass == 0 && ms > 0.f) iters = (int)std::min(4.0e6, std::max(1000.0, iters * (double)ms_target / ms));
    }
    *dmma_tflops = best;
    return SGP_OK;
}
