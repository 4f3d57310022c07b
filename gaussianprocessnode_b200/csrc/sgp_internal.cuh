// libsgp internals shared by the translation units (not part of the public ABI; see include/sgp.h).
#pragma once
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>
#include "../../include/sgp.h"

#define SGP_MAX_D 16
#define SGP_EXP_TAB 2048                       // entries of the 2^(j/2048) table
#define SGP_EXP_SCALE 2954.639443740597       // 2048 / ln 2

struct SgpComm;                                // NCCL handle (comm.cu)

struct sgp_ctx {
    int dev = 0;
    int num_sms = 148;
    cudaStream_t stream = nullptr;
    cudaStream_t stream2 = nullptr;           // copies that overlap the next kernel (created with the context)
    cudaEvent_t ev_copy = nullptr;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    std::string err;

    // kernel(theta)
    int kind = SGP_KERNEL_SE, D = 0;
    double variance = 1.0;
    double ell[SGP_MAX_D] = {0};
    bool have_kernel = false;

    // inducing inputs
    int M = 0;
    double* Z_dev = nullptr;        // M x D raw (point-major)
    double center[SGP_MAX_D] = {0}; // mean of Z: both X and Z are shifted by it before scaling (shift invariance)
    bool have_Z = false;

    // data (resident)
    int64_t N = 0, Ncap = 0;        // Ncap: allocated points (multiple of the chunk size, zero padded)
    int Dcap = 0;                   // input dimension the owned X buffer was allocated for
    double *X_dev = nullptr, *y_dev = nullptr, *yv_dev = nullptr, *w_dev = nullptr;
    bool own_data = false, have_yv = false, have_w = false, have_data = false;     // have_data: sgp_set_data[_dev] was called (N may be 0)

    // statistics of the last sweep: stats_dev = [psi2 (M*M) | psi1 (M*Dout) | psi0 | sum_y2 | sum_w | n]
    double* stats_dev = nullptr;
    size_t stats_cap = 0;
    int Dout = 1;
    bool have_stats = false;
    bool stats_of_data = false;     // ... and they are the statistics of the resident data set (not of a sigma-point cloud / theta scratch)

    // scratch
    double* work_dev = nullptr;  size_t work_cap = 0;      // split-N partials
    double* zrec_dev = nullptr;  size_t zrec_cap = 0;      // prepared inducing rows
    void* kbuf_window = nullptr; size_t kbuf_window_bytes = 0;   // L2 access-policy window currently set on the stream
    double* kbuf_dev = nullptr;  size_t kbuf_cap = 0;      // L2-resident K_uf panel of one slab (generate-once sweep)
    double* fetch_host = nullptr; size_t fetch_cap = 0;    // pinned staging of the small results (Psi1 | scalars)
    // sgp_sweep_psi_host: the upload runs on stream2 BESIDE the sweep kernel's launch; the kernel waits for `ready_dev` (a word the copy engine
    // writes after the data, in stream order) before it touches the data.  upload_pending: copies are in flight on stream2 that the main stream
    // has not been ordered after yet (sgp_join_upload orders it; the generate-once sweep consumes the flag instead).
    unsigned* ready_dev = nullptr; unsigned* ready_host = nullptr; unsigned ready_epoch = 0; bool upload_pending = false;
    // packed lower triangle of Psi2 (LAPACK 'L' packed storage) for the host: written by the sweep's phase 2 / the exchange, or by pack_lower_kernel
    double* packed_dev = nullptr; size_t packed_cap = 0;
    bool want_packed = false;                              // the running call returns Psi2 packed: the sweep should leave it in packed_src
    const double* packed_src = nullptr;                    // ... where the last sweep left it (nullptr: not produced, fetch packs it)
    void* flush_dev = nullptr; size_t flush_cap = 0;       // L2 flush buffer of sgp_sweep_timed_flushed
    unsigned sweep_bar_epoch = 0;                          // launch counter of the plain-launch grid barrier (SGP_SWEEP_COOP=0)
    unsigned* sweep_flags_dev = nullptr;                   // generation / consumption counters of the generate-once sweep
    int* p2plan_dev = nullptr; size_t p2plan_cap = 0;      // phase-2 plan of the generate-once sweep (sweep.cu: build_p2_plan) ...
    long long p2plan_key[8] = {-1, -1, -1, -1, -1, -1, -1, -1};   // ... and the configuration it was built for
    int p2plan_off[4] = {0, 0, 0, 0};                      // int offsets of {cta_off, items, tile_off, slots} inside p2plan_dev
    double* exptab_dev = nullptr;
    double* dense_dev = nullptr; size_t dense_cap = 0;     // M x M scratch for factorisations
    int* info_dev = nullptr;
    double* pred_dev = nullptr; size_t pred_cap = 0;       // scratch of sgp_predict_mean / sgp_predict_probit
    double* wt_dev = nullptr;                              // block partials + ticket of sgp_wterms_reduce
    double* in_dev = nullptr; size_t in_cap = 0;           // scratch of sgp_in_logmessage (inmsg.cu)
    double* theta_dev = nullptr; size_t theta_cap = 0;     // scratch of the theta objective / gradient (theta.cu)

    // K_uu factor
    double* KuuL_dev = nullptr; int KuuL_M = 0; bool have_kuu = false; double kuu_jitter = 0.0;
    bool kuu_jitter_known = false, kuu_speculative = false; int info2_last = 0;    // K_uu job sharing the posterior's launch (api.cu posterior_core)
    double* Kinv_dev = nullptr;                            // K_uu^-1 (full symmetric), refreshed by sgp_kuu_factor
    double* kuu_dinv_dev = nullptr; size_t kuu_dinv_cap = 0; // inverses of the 64 x 64 diagonal blocks of KuuL
    double* dinv_dev = nullptr; size_t dinv_cap = 0;       // ... of the factor produced by the last sgp_potrf_lower

    // uncertain-input scratch (sigma-point cloud)
    double *sp_X_dev = nullptr, *sp_w_dev = nullptr, *sp_y_dev = nullptr; size_t sp_cap = 0;
    int64_t sp_N = 0; int sp_S = 0;                         // nodes / sigma points per node of the resident cloud (0 = none)
    double* unc_dev = nullptr; size_t unc_cap = 0;          // arena for the per-call scratch of sgp_sweep_psi_uncertain

    // resident prior / posterior of v (api.cu: post_*): [Lambda_prior | xi_prior | mu | Sigma | Uv]
    double* post_dev = nullptr; size_t post_cap = 0; int post_M = 0;
    bool have_prior = false, have_post = false, have_post_uv = false;

    SgpComm* comm = nullptr;

    // last sweep launch record
    int last_launches = 0, last_grid = 0, last_block = 0, last_smem = 0;
    bool want_exchange = false;           // set by the public sweep calls: theta / uncertain-input sweeps keep their statistics local
    bool last_sweep_exchanged = false;    // the last sweep kernel already summed the statistics over the ranks
    float last_main_ms = 0.f;
    bool dense_timing = false;            // sgp_dense_timed: the M x M calls record ev[1] after their last kernel
    long long* sweep_dbg_dev = nullptr;   // optional per-segment clocks (sgp_sweep_debug_clocks)
    int sweep_dbg_slots = 0;
};

#define SGP_CUDA(ctx, call)                                                                          \
    do {                                                                                             \
        cudaError_t e__ = (call);                                                                    \
        if (e__ != cudaSuccess) {                                                                    \
            (ctx)->err = std::string(#call) + ": " + cudaGetErrorString(e__);                        \
            return SGP_ERR_CUDA;                                                                     \
        }                                                                                            \
    } while (0)

// NVTX range over a host-side call of the library (upload / sweep / posterior / fetch): visible in Nsight Systems timelines, free otherwise
struct SgpRange {
    explicit SgpRange(const char* name) { nvtxRangePushA(name); }
    ~SgpRange() { nvtxRangePop(); }
};
#define SGP_RANGE(name) SgpRange sgp_range__(name)

#define SGP_FAIL(ctx, code, msg) do { (ctx)->err = (msg); return (code); } while (0)

int sgp_ensure(sgp_ctx* ctx, double** p, size_t* cap, size_t need_doubles);
int sgp_ensure_zero(sgp_ctx* ctx, double** p, size_t* cap, size_t need_doubles);      // zero-filled when (re)allocated

int sgp_join_upload(sgp_ctx* ctx);     // api.cu: orders the main stream after an upload still in flight on the copy stream (no-op otherwise)
// sweep.cu
int sgp_sweep_launch(sgp_ctx* ctx, const double* X, const double* y, const double* yv, const double* w, int64_t N,
                     int64_t Ncap, bool time_main);
// dense_coop.cu: one cooperative kernel per M x M job: [build A] -> Cholesky -> [X = L^-1, S = X'X] -> [mu = S xi] -> [Ut = L']
constexpr int SGP_FLIP_UV_MAX_M = 4096;      // four M-vectors of the closed-form rank-one factor live in shared memory
struct SgpDenseJob {
    int M = 0;
    int build = 0;             // 0: A as given | 1: A = P + w S2, xi = xip + w s1 (carry: P <- A, xip <- xi) | 2: A = K_uu(Z) + jitter I | 3: A = Sig + mu_in mu_in'
    double* A = nullptr;       // M x M column-major, factored in place (L lower, strict upper triangle zeroed)
    double* Dinv = nullptr;    // [ceil(M/64)][64 x 64] inverses of the diagonal blocks of L
    const double* S2 = nullptr; const double* s1 = nullptr; double* P = nullptr; double* xip = nullptr; double* xi = nullptr; double w = 0.0; int carry = 0;
    double jitter = 0.0;
    const double* Sig = nullptr; const double* mu_in = nullptr;
    double* X = nullptr; double* Tmp = nullptr; double* S = nullptr;    // optional inverse: X = L^-1, S = (L L')^-1 (full symmetric), Tmp = M x M scratch
    double* mu = nullptr;      // optional: mu = S xi
    double* Ut = nullptr;      // optional: Ut = L' (upper)
    int flip = 0;              // build 1 with X and S: factorise the index-reversed matrix; S is scratch, Sout / mu / Uv come out un-reversed
    double* Sout = nullptr;    // flip: A^-1 (full symmetric)
    double* Uv = nullptr;      // flip, optional (needs mu, M <= SGP_FLIP_UV_MAX_M): chol(A^-1 + mu mu').U from a running sum down the columns of X
    long long* clk = nullptr;  // optional: 8 phase clocks of CTA 0
    int reset_info = 1;        // zero ctx->info_dev before the launch
};
int sgp_dense_job(sgp_ctx* ctx, const SgpDenseJob& job, const SgpDenseJob* second = nullptr);   // enqueues (optionally a second independent job of the same M in the same launch; its pivot record: info_dev[1]); no host synchronisation
// pinned host staging for the small read-backs (ctx->fetch_host), at least n doubles; nullptr on failure
inline double* sgp_host_stage(sgp_ctx* ctx, size_t n) {
    if (ctx->fetch_cap < n) {
        if (ctx->fetch_host) cudaFreeHost(ctx->fetch_host);
        ctx->fetch_host = nullptr; ctx->fetch_cap = 0;
        if (cudaHostAlloc((void**)&ctx->fetch_host, (n + 1024) * sizeof(double), cudaHostAllocDefault) != cudaSuccess) return nullptr;
        ctx->fetch_cap = n + 1024;
    }
    return ctx->fetch_host;
}
int sgp_kuu_factor_enqueue(sgp_ctx* ctx, double jitter);                         // api.cu: K_uu job without the host synchronisation
int sgp_dense_info(sgp_ctx* ctx, const char* what);                               // synchronises; non-positive pivot -> SGP_ERR_NOT_PD
int sgp_potrf_lower(sgp_ctx* ctx, double* A, int M);                              // in place, column-major, lower (synchronises)
int sgp_trsm_lower_dinv(sgp_ctx* ctx, const double* L, const double* dinv, double* B, double* tmp, int M, int nrhs, bool trans);
// dense.cu: deterministic reductions
int sgp_dot(sgp_ctx* ctx, const double* a, size_t sa, const double* b, size_t sb, size_t n, double* out);
// out[0] = <Kinv, Psi2>, out[1] = <R + mu_outer mu_outer', Psi2>, out[2] = mu' psi1, out[3] = trace(Kinv)   (R, mu_outer, Kinv may be null)
int sgp_wterms_reduce(sgp_ctx* ctx, const double* Kinv, const double* Psi2, const double* R, const double* mu_outer, const double* mu,
                      const double* psi1, int M, double* out);
// api.cu: resident posterior (mu [M], Uv [M x M upper]) or nullptr when there is none
const double* sgp_resident_mu(sgp_ctx* ctx);
const double* sgp_resident_uv(sgp_ctx* ctx);
const double* sgp_resident_sigma(sgp_ctx* ctx);
int sgp_sweep_resident(sgp_ctx* ctx, bool time_main);     // sweep of the resident data, statistics summed over the ranks
// comm.cu
int sgp_comm_allreduce(sgp_ctx* ctx, double* buf, size_t count);          // xchg.cu: peer-memory kernel, or NCCL when the regions are not mapped
int sgp_comm_allreduce_stats(sgp_ctx* ctx, int M, int D_out);
int sgp_comm_barrier(sgp_ctx* ctx);                                         // flag barrier over the ranks on the ctx stream (peer-memory exchange only)               // ... of the resident statistics (packed lower triangle on the wire)
void sgp_comm_destroy(sgp_ctx* ctx);
// Peer-memory exchange (xchg.cuh; single node, <= 8 ranks): every rank owns a region [flags (256 B) | contribution | result] (cap doubles
// each) that all peers map through CUDA IPC.
struct SgpXchg {
    int nranks = 1, rank = 0;
    unsigned epoch = 0;            // barrier value of this exchange (flags are monotonic, never reset)
    char* peers[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};   // region of rank q as mapped here
    size_t slot0_off = 0, slot_bytes = 0;   // byte offset of the contribution buffer inside a region; the result buffer follows after slot_bytes
};
bool sgp_comm_xchg(sgp_ctx* ctx, size_t need_doubles, SgpXchg* x);     // fills x (and bumps the epoch) if the peer-memory exchange is available
inline int sgp_ensure_stats(sgp_ctx* ctx, size_t need_doubles) { return sgp_ensure(ctx, &ctx->stats_dev, &ctx->stats_cap, need_doubles); }

// ---------------------------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// exp(t) for an argument ALREADY scaled by 2048/ln2 (tp = t * SGP_EXP_SCALE).  Range reduction by the magic-number
// trick, 2^(j/2048) from a shared-memory table, cubic Taylor remainder (|x| <= ln2/4096: x^4/24 < 4e-17), exponent
// patched with integer adds.  7 FP64-pipe instructions.  Returns 0 below exp(-677).
__device__ __forceinline__ double exp_scaled(double tp, const double* __restrict__ tab) {
    const double MAGIC = 6755399441055744.0;            // 1.5 * 2^52
    const double C1 = 3.384507717577858e-04;           // ln2/2048
    const double C2 = 5.72744624517204e-08;           // C1^2 / 2
    const double C3 = 6.461528672932365e-12;           // C1^3 / 6
    double m = tp + MAGIC;
    int n = __double2loint(m);
    double nd = m - MAGIC;
    double r = tp - nd;
    double p = fma(r, C3, C2);
    p = fma(p, r, C1);
    p = p * r;
    double T = tab[n & (SGP_EXP_TAB - 1)];
    double res = fma(T, p, T);
    int hi = __double2hiint(res) + ((n >> 11) << 20);
    res = __hiloint2double(hi, __double2loint(res));
    // tp < -2.0e6  <=>  sign set and magnitude above: compare the high word as an unsigned integer (ALU pipe)
    return ((unsigned)__double2hiint(tp) > 0xC13E8480u) ? 0.0 : res;   // hi word of -2.0e6 = 0xC13E8480
}
// U interleaved evaluations of exp_scaled (independent chains written stage by stage so that ptxas interleaves them).
template <int U>
__device__ __forceinline__ void exp_scaled_v(const double (&tp)[U], double (&out)[U], const double* __restrict__ tab) {
    const double MAGIC = 6755399441055744.0;
    const double C1 = 3.384507717577858e-04, C2 = 5.72744624517204e-08, C3 = 6.461528672932365e-12;
    double m[U], r[U], p[U], T[U];
    int n[U];
#pragma unroll
    for (int u = 0; u < U; ++u) m[u] = tp[u] + MAGIC;
#pragma unroll
    for (int u = 0; u < U; ++u) { n[u] = __double2loint(m[u]); T[u] = tab[n[u] & (SGP_EXP_TAB - 1)]; }
#pragma unroll
    for (int u = 0; u < U; ++u) m[u] = m[u] - MAGIC;
#pragma unroll
    for (int u = 0; u < U; ++u) r[u] = tp[u] - m[u];
#pragma unroll
    for (int u = 0; u < U; ++u) p[u] = fma(r[u], C3, C2);
#pragma unroll
    for (int u = 0; u < U; ++u) p[u] = fma(p[u], r[u], C1);
#pragma unroll
    for (int u = 0; u < U; ++u) p[u] = p[u] * r[u];
#pragma unroll
    for (int u = 0; u < U; ++u) {
        double res = fma(T[u], p[u], T[u]);
        int hi = __double2hiint(res) + ((n[u] >> 11) << 20);
        res = __hiloint2double(hi, __double2loint(res));
        out[u] = ((unsigned)__double2hiint(tp[u]) > 0xC13E8480u) ? 0.0 : res;
    }
}
#endif
