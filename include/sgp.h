/*
 * libsgp -- C ABI of the B200-native data sweep for sparse variational GP factor nodes.
 *
 * Drop-in boundary for ONE hot path of biaslab/GaussianProcessNode: the per-iteration pass over N inputs that
 * builds K_uf and accumulates Psi0 / Psi1 / Psi2, plus the M x M K_uu / posterior factorisations.  The reference
 * has no FFI: the path sits behind ReactiveMP's @rule / @average_energy dispatch (Julia multiple dispatch).  Each
 * entry point below states the reference code it replaces (file:line under /root/reference); INTEGRATION.md shows
 * the `ccall` stubs a maintainer adds to GPnode/UniSGPnode.jl / MultiSGPnode.jl so the rule surface stays unchanged.
 *
 * Conventions
 *   - plain C, no torch / CUDA types in any signature; `int` return: 0 = ok, < 0 = error (see SGP_ERR_*),
 *     human-readable text via sgp_last_error(ctx).
 *   - every pointer is a HOST pointer owned by the caller unless the name ends in `_dev`; the library owns all
 *     device memory.  Matrices are column-major Float64, exactly Julia's layout: X is D x N (point n = D
 *     contiguous doubles), Z is D x M, Psi2 / L / Sigma are M x M.
 *   - one sgp_ctx per GPU per host task; a ctx is NOT thread-safe.  All work of a ctx runs on its own stream.
 *   - there is no CPU fallback: without a CUDA device every call fails with SGP_ERR_CUDA.
 */
#ifndef SGP_H
#define SGP_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct sgp_ctx sgp_ctx;

enum {
    SGP_OK = 0,
    SGP_ERR_ARG = -1,        /* bad argument / call order */
    SGP_ERR_CUDA = -2,       /* CUDA runtime failure (no device, OOM, launch failure) */
    SGP_ERR_NOT_PD = -3,     /* Cholesky met a non-positive pivot (reference: FastCholesky falls back silently) */
    SGP_ERR_UNSUPPORTED = -4,
    SGP_ERR_COMM = -5        /* NCCL failure */
};

/* kernel kinds: KernelFunctions convention, r = ||(x - z) ./ ell|| (experiments/regression_kin40k.ipynb:108) */
enum { SGP_KERNEL_SE = 0, SGP_KERNEL_MATERN32 = 1, SGP_KERNEL_MATERN52 = 2 };

/* expectation methods for uncertain inputs (ReactiveMP srcubature()/ghcubature(p), helper_functions/ut_approx.jl,
 * closed-form SE-ARD extension) */
enum { SGP_METHOD_SRCUBATURE = 0, SGP_METHOD_GENUT = 1, SGP_METHOD_GAUSSHERMITE = 2, SGP_METHOD_CLOSED_FORM_SE = 3,
       SGP_METHOD_POINT = 4 /* q(x_n) = PointMass(mean_n): one point of weight 1, `cov` may be NULL -- gives the per-node terms of
                               sgp_uncertain_node_terms to the `q_in::PointMass` rules that clamp per node (GPnode/UniSGPnode.jl:438-458) */ };

/* ---- lifetime -------------------------------------------------------------------------------------------- */
int sgp_create(sgp_ctx** ctx, int device_id);
void sgp_destroy(sgp_ctx* ctx);
const char* sgp_last_error(const sgp_ctx* ctx);
/* library build string: "libsgp <version> sm_100a" */
const char* sgp_version(void);

/* Page-locked host memory for the arrays that cross the boundary every sweep (X, ybar, the Psi outputs): copies from / to
 * such buffers run at full PCIe speed and without a staging copy.  Any other host memory is accepted everywhere too.
 * (Julia: wrap with unsafe_wrap(Array, Ptr{Float64}(p), dims); free after the last use.) */
int sgp_pinned_alloc(size_t bytes, void** ptr);
void sgp_pinned_free(void* ptr);

/* ---- model state (what UniSGPMeta / MultiSGPMeta carry: helper_functions/gp_helperfunction.jl:33-73) ------ */
/* kernel(theta) of the meta, already transformed by the host (softplus etc.): variance sigma^2, lengthscale[D]. */
int sgp_set_kernel(sgp_ctx* ctx, int kind, int D, double variance, const double* lengthscale);
/* meta.Xu flattened to D x M. */
int sgp_set_inducing(sgp_ctx* ctx, int M, const double* Z);
/* The data the N per-point rules would each see one element of: inputs X (D x N), targets ybar (y_n, or E[f_n] for
 * classification; NULL = zeros), yvar (Var[f_n]; NULL = zeros), per-point weights wts (NULL = ones; sigma-point
 * weights may be negative).  Copied to the device once and kept resident across VMP iterations. */
int sgp_set_data(sgp_ctx* ctx, int64_t N, const double* X, const double* ybar, const double* yvar, const double* wts);
/* replace only the targets (classification: E[f_n], Var[f_n] change every VMP iteration, X does not) */
int sgp_set_targets(sgp_ctx* ctx, const double* ybar, const double* yvar);
/* same as sgp_set_data but the pointers are device pointers on this ctx's device; data is used in place. */
int sgp_set_data_dev(sgp_ctx* ctx, int64_t N, const double* X_dev, const double* ybar_dev, const double* yvar_dev,
                     const double* wts_dev);

/* ---- the sweep ------------------------------------------------------------------------------------------- */
/* Replaces N invocations of `@rule UniSGP(:v)` (GPnode/UniSGPnode.jl:144-158, 161-173) and the running sums of
 * `prod` (:62-73), and supplies what N invocations of `@rule UniSGP(:w)` (:196-238) need:
 *   psi0 = sum_n w_n k(x_n,x_n),  psi1[M] = K_uf (w .* ybar),  psi2[M*M] = K_uf diag(w) K_uf' (full symmetric),
 *   sum_y2 = sum_n w_n (ybar_n^2 + yvar_n).
 * Any output pointer may be NULL (result stays on the device for the calls below).  With a communicator attached
 * (sgp_comm_init) the call is COLLECTIVE: every rank must make it, and the statistics are summed over the ranks before they are
 * returned (inside the sweep kernel through NVLink peer memory, or by one NCCL all-reduce); a rank that never arrives makes the
 * others fail with SGP_ERR_CUDA after a time-out instead of hanging. */
int sgp_sweep_psi(sgp_ctx* ctx, double* psi0, double* psi1, double* psi2, double* sum_y2);

/* sgp_set_data + sgp_sweep_psi as ONE call with a single host synchronisation: the per-step call of a host whose data change
 * every step (mini-batches).  The data stay resident afterwards exactly as after sgp_set_data.  The upload runs on the library's copy
 * stream BESIDE the launch of the sweep kernel, which waits on the device for a word the copy engine writes after the data (M > 384;
 * SGP_HOST_OVERLAP=0 restores upload-then-launch); the host buffers must stay untouched until the call returns, as before. */
int sgp_sweep_psi_host(sgp_ctx* ctx, int64_t N, const double* X, const double* ybar, const double* yvar, const double* wts,
                       double* psi0, double* psi1, double* psi2, double* sum_y2);
/* The same call with ALL statistics returned in ONE packed buffer -- the layout the multi-GPU exchange uses --:
 *   stats_packed = [ lower triangle of Psi2, column by column: stats_packed[i + j (2M - j - 1) / 2] = Psi2[i, j], i >= j  (M (M + 1) / 2 doubles: LAPACK's packed
 *                    storage for uplo = 'L'; dpptrf / dspmv take it as is)  |  Psi1 (M)  |  Psi0, sum_y2, sum_w, n ]
 * i.e. M (M + 1) / 2 + M + 4 doubles, filled by ONE device-to-host copy.  The running sum of `prod` (GPnode/UniSGPnode.jl:62-73) is symmetric, so this is all of
 * it at half the bytes over the bus; the sweep's last phase writes the packed copy itself (no extra pass, no extra launch).  psi0 / psi1 / sum_y2 are optional
 * conveniences (copied out of the tail of stats_packed on the host; NULL to skip). */
int sgp_sweep_psi_host_packed(sgp_ctx* ctx, int64_t N, const double* X, const double* ybar, const double* yvar, const double* wts,
                              double* psi0, double* psi1, double* stats_packed, double* sum_y2);
/* Psi2 of the last sweep (any of the sweeps above) as its packed lower triangle alone: M (M + 1) / 2 doubles, laid out as in sgp_sweep_psi_host_packed. */
int sgp_fetch_psi2_packed(sgp_ctx* ctx, double* psi2_packed);

/* Uncertain inputs q(x_n) = N(mean_n, cov_n): replaces the cubature loop `approximate_kernel_expectation(!)`
 * (GPnode/UniSGPnode.jl:11-37, GPnode/MultiSGPnode.jl:11-35) inside `@rule UniSGP(:v)` (:125-140) and
 * `@rule MultiSGP(:v)/(:w)/(:out)` (GPnode/MultiSGPnode.jl:290-328, 367-444, 90-120), summed over the N nodes.
 *   mean: d x N, cov: d x d x N (column-major per point), p: Gauss-Hermite order (method 2 only),
 *   R: N x D_out row weights (MultiSGP: Y * W_bar, row n = (W_bar' mu_y_n)'; UniSGP: ybar; NULL = ones, D_out = 1)
 *   psi1[M*D_out] = sum_n Psi1_n r_n' (column d = output d), psi2[M*M] = sum_n Psi2_n (no jitter added),
 *   psi1_n[M*N] (optional, may be NULL) = per-point Psi1_n, needed by the :out / :w rules. */
int sgp_sweep_psi_uncertain(sgp_ctx* ctx, int method, int p, int64_t N, const double* mean, const double* cov,
                            int D_out, const double* R, double* psi0, double* psi1, double* psi2, double* psi1_n);

/* ---- M x M factorisations --------------------------------------------------------------------------------- */
/* K_uu = kernelmatrix(kernel(theta), Xu) + jitter*I ; L = fastcholesky!(K_uu).L
 * (experiments/regression_kin40k.ipynb:183-184, classification_banana.ipynb:163-164).  L (M x M, lower, may be NULL). */
int sgp_kuu_factor(sgp_ctx* ctx, double jitter, double* L);
/* B <- K_uu^{-1} B for nrhs right-hand sides (M x nrhs); cholinv(Kuu) = solve with B = I
 * (experiments/Pendulum_Wishart_2d.ipynb:2542-2543). */
int sgp_kuu_solve(sgp_ctx* ctx, int nrhs, double* B);
/* The N-th `prod` (GPnode/UniSGPnode.jl:62-73) on the statistics of the last sweep:
 *   Lambda = Lambda0 + w Psi2, xi = xi0 + w Psi1, Sigma_v = cholinv(Lambda), mu_v = Sigma_v xi,
 *   Uv = fastcholesky!(Sigma_v + mu_v mu_v').U.   Lambda0 (M x M) / xi0 (M) are the prior's natural parameters.
 * Outputs may be NULL.  ONE kernel launch (M <= 4096): Lambda is factorised in reversed index order, so that the inverse of the factor is the
 * Cholesky factor of Sigma_v and Uv follows by a closed-form rank-one update; a non-positive pivot is reported as SGP_ERR_NOT_PD (the row
 * number in the message counts from the END of the matrix).  When K_uu is stale (sgp_set_kernel / sgp_set_inducing since the last
 * sgp_kuu_factor) and a jitter is on record, K_uu is refactored with that jitter in the same launch (as if sgp_kuu_factor had been
 * called; a failure of that part is not this call's error -- the next sgp_kuu_factor reports it). */
int sgp_posterior_v(sgp_ctx* ctx, const double* xi0, const double* Lambda0, double w, double* mu_v, double* Sigma_v,
                    double* Uv);
/* Streaming prior (experiments/regression_kin40k.ipynb:200-213: the posterior of mini-batch b is the prior of mini-batch b+1,
 * restarted from N(0, 50 I) every epoch): the Gaussian on v stays RESIDENT on the device.
 *   sgp_prior_set            load a prior (host natural parameters xi0 [M], Lambda0 [M x M])
 *   sgp_prior_set_isotropic  N(0, variance I) without any host traffic
 *   sgp_posterior_v_stream   the N-th prod on the resident prior and the last sweep; carry != 0: the posterior's natural
 *                            parameters become the resident prior.  Outputs may be NULL (nothing is copied back); mu_v / Sigma_v
 *                            of the last posterior stay resident for sgp_w_terms / sgp_theta_objective (pass NULL there).
 *                            Uv == NULL: the factor of Sigma_v + mu_v mu_v' is not formed; the resident consumers use
 *                            <R_v, Psi2> = <Sigma_v, Psi2> + mu_v' Psi2 mu_v instead. */
int sgp_prior_set(sgp_ctx* ctx, const double* xi0, const double* Lambda0);
int sgp_prior_set_isotropic(sgp_ctx* ctx, double variance);
int sgp_posterior_v_stream(sgp_ctx* ctx, double w, int carry, double* mu_v, double* Sigma_v, double* Uv);
/* sum over n of the :w rule / energy ingredients (GPnode/UniSGPnode.jl:196-238, 337-387), from the last sweep:
 *   sumI1 = Psi0 - tr(K_uu^{-1} Psi2),  sumI2 = sum_y2 - 2 mu_v' Psi1 + <Uv' Uv, Psi2>.
 * mu_v / Uv are INPUTS (the previous sweep's posterior, as the VMP schedule has it), or both NULL = the resident posterior of
 * the last sgp_posterior_v[_stream]; needs sgp_kuu_factor. */
int sgp_w_terms(sgp_ctx* ctx, const double* mu_v, const double* Uv, double* sumI1, double* sumI2);
/* `@rule UniSGP(:out)` over a test set (GPnode/UniSGPnode.jl:96-104; regression_kin40k.ipynb:289-304): out = K_*u mu_v */
int sgp_predict_mean(sgp_ctx* ctx, int64_t Nt, const double* Xt, const double* mu_v, double* out);

/* `predict_new` of the classification drivers (experiments/classification_banana.ipynb:289-305, GPT_classification.ipynb cell 13): the
 * `@rule UniSGP(:out)` message N(mean_f[n], 1 / w_bar) (GPnode/UniSGPnode.jl:96-104) pushed through ReactiveMP's `@rule Probit(:out)`:
 *   mean_f[n] = K_*u mu_v,  *var_f = 1 / w_bar (the :out message's variance: the same for every test point),
 *   prob_y[n] = Phi(mean_f[n] / sqrt(1 + 1 / w_bar))   (mean of the Bernoulli; its variance is p (1 - p)).  var_f may be NULL. */
int sgp_predict_probit(sgp_ctx* ctx, int64_t Nt, const double* Xt, const double* mu_v, double w_bar, double* mean_f, double* var_f,
                       double* prob_y);

/* ---- the theta step ---------------------------------------------------------- */
/* Collapsed objective of the hyper-parameter step and its exact gradient on the resident data, at the kernel parameters
 * currently set: replaces `neg_log_backwardmess_fast` + `ForwardDiff.gradient!`
 * (helper_functions/derivative_helper.jl:23-39, 55-67; experiments/regression_kin40k.ipynb:214-222):
 *   F = sum_n [ w/2 k_nn - w/2 |L_u^-1 k_n|^2 + w/2 |Uv k_n|^2 - w y_n mu_v' k_n ],   K_uu = L_u L_u' (+ jitter I)
 * value = F, dvariance = dF/d sigma^2, dlengthscale[D] = dF/d ell_d (analytic; the host applies its own chain rule for the
 * raw parameters, e.g. softplus').  mu_v (M) and Uv (M x M upper, column-major) are inputs (both NULL = resident posterior).  Any output may be NULL; without
 * gradient outputs only the value is computed.  The statistics of the last sgp_sweep_psi on the same data and kernel are reused (otherwise the sweep
 * runs first); K_uu is refactored only when the kernel, Z or the jitter changed.  The whole step is enqueued without intermediate host
 * synchronisation and ends in one read-back.  With a communicator attached the call is COLLECTIVE: the
 * statistics are the rank-summed ones and the rank-local part of the gradient is summed over the ranks. */
int sgp_theta_objective(sgp_ctx* ctx, const double* mu_v, const double* Uv, double w, double jitter, double* value,
                        double* dvariance, double* dlengthscale);

/* Per-NODE expectations for the uncertain-input UniSGP rules, from the sigma-point cloud of the preceding sgp_sweep_psi_uncertain (methods 0-2,
 * resident on the device): `@rule UniSGP(:w)` (GPnode/UniSGPnode.jl:177-192), `:out` (:85-93), `@average_energy` (:290-313).  These rules clamp
 * I1 and I2 per node, so they need per-node values, not sums:
 *   psi0_n = Psi0_n,  q_kinv_n = tr(Kuu^-1 Psi2_n),  lin_n = Psi1_n' mu_v (the :out mean),  q_rv_n = tr(Uv' Uv Psi2_n)   -- N doubles each, any may be NULL;
 *   tr_kinv = tr(Kuu^-1), frob_uv = |Uv|_F^2 (what the reference's `+ 1e-8 I` on Psi2 adds to the two traces).
 * The host finishes I1 = clamp(psi0_n - q_kinv_n - 1e-8 tr_kinv), I2 = clamp(y^2 + v - 2 y lin_n + q_rv_n + 1e-8 frob_uv).  mu_v (M), Uv (M x M upper,
 * column-major) are inputs; needs sgp_kuu_factor. */
int sgp_uncertain_node_terms(sgp_ctx* ctx, const double* mu_v, const double* Uv, double* psi0_n, double* q_kinv_n, double* lin_n, double* q_rv_n,
                             double* tr_kinv, double* frob_uv);

/* ---- backward message towards the input of MultiSGP ------------------------------------- */
/* `@rule MultiSGP(:in, Marginalisation)` (GPnode/MultiSGPnode.jl:162-185, 187-211, 213-236) for N nodes at once.  Per node the reference
 * returns the closure
 *   log_backwardmess(x) = -1/2 tr(W) (k(x,x) - k_u(x)' Kuu^-1 k_u(x)) + sumdiagV' k_u(x) - 1/2 k_u(x)' sumRvblk_W k_u(x),
 * sumdiagV = sum_d mu_v^(d) (W mu_y)_d, sumRvblk_W = sum_ij W_ij R_v^{ij}, which ReactiveMP evaluates at the cubature points of the forward
 * message (`prod` override :38-45) or minimises (LBFGS, :213-236).  This call evaluates it at P points for each of the N nodes:
 *   Xp: d x P x N (point p of node n = d contiguous doubles),  R: N x D_out, row n = (W mu_y,n)' (NULL = ones),  Mv: M x D_out, column d =
 *   mu_v^(d),  S: M x M = sumRvblk_W,  trW = tr(W);   f: P x N values,  grad: d x P x N (optional),  hess: d x d x P x N (optional; analytic,
 *   SE-ARD only -- what ForwardDiff / Zygote.hessian return in the Laplace variant).  Needs sgp_kuu_factor (K_uu^-1). */
int sgp_in_logmessage(sgp_ctx* ctx, int64_t N, int P, const double* Xp, int D_out, const double* R, const double* Mv, const double* S,
                      double trW, double* f, double* grad, double* hess);

/* ---- multi-GPU: N is sharded over ranks, one ctx per rank/GPU --------------------------------------------- */
/* NCCL unique id (128 bytes) created on rank 0 and handed to the other ranks by the host.  sgp_comm_init is collective: it creates the
 * NCCL communicator and, on a single node with peer access (<= 8 ranks), maps one exchange region per rank into all peers through CUDA
 * IPC so that the sweep kernel can sum the statistics itself (SGP_COMM_P2P=0 forces the NCCL path; M <= 2048 by default,
 * SGP_COMM_P2P_MAXM raises it). */
int sgp_comm_unique_id(char id[128]);
int sgp_comm_init(sgp_ctx* ctx, int nranks, int rank, const char id[128]);

/* ---- measurement hooks (bench.py) ------------------------------------------------------------------------ */
/* Runs the sweep `reps` times on resident data without host copies and returns the mean device time of one sweep
 * (CUDA events on the ctx stream) and of its dominant kernel alone. */
int sgp_sweep_timed(sgp_ctx* ctx, int reps, float* ms_per_sweep, float* ms_main_kernel);
/* The same with an L2 flush (a `flush_mb` MB device buffer is rewritten on the ctx stream) before every repetition, all repetitions
 * enqueued without a host synchronisation; the flush is outside the timed intervals (one event pair per repetition).  With a
 * communicator attached the ranks are aligned by a flag barrier on the stream between the flush and the start of every timed interval,
 * and stay in lock step through the exchange itself.  ms_main_kernel == NULL: the main kernel is not bracketed by its own event pair (the
 * timed interval then contains the sweep alone). */
int sgp_sweep_timed_flushed(sgp_ctx* ctx, int reps, int flush_mb, float* ms_per_sweep, float* ms_main_kernel);
/* Device time of the M x M entry points on the resident state (CUDA events on the ctx stream around the call's kernels; no host copies):
 * what = 0: sgp_kuu_factor(jitter), 1: sgp_posterior_v_stream(w, carry = 0) with Uv, 2: the same without Uv, 3: sgp_w_terms(NULL, NULL).
 * Mean over `reps` calls. */
int sgp_dense_timed(sgp_ctx* ctx, int what, double w, double jitter, int reps, float* ms_per_call);
/* FP64 tensor-pipe (DMMA.8x8x4) peak of THIS device in TFLOP/s: a register-only loop of independent accumulator chains, about
 * `ms_target` milliseconds per pass, best of three.  The denominator of the sweep's roofline fraction (MEASURED_PEAKS.json has no FP64 entry). */
int sgp_fp64_peak(sgp_ctx* ctx, int ms_target, double* dmma_tflops);
/* number of kernels the last sweep launched, and the main kernel's launch geometry */
int sgp_last_sweep_info(sgp_ctx* ctx, int* n_launches, int* grid, int* block, int* smem_bytes);
/* per-segment clock counters of the fused sweep kernel (load-balance tuning).  The first call switches the
 * recording on; later calls copy out up to `cap` records {chunks, SM clocks, is_diagonal_tile, cta} (-1 = unused slot)
 * of the last sweep and return their number in *nrec. */
int sgp_sweep_debug_clocks(sgp_ctx* ctx, int64_t* out, int cap, int* nrec);
/* The plan of the generate-once sweep's last phase (deterministic reduction of the segment partials) for a configuration, as the kernel gets it -- pure host
 * arithmetic, no device needed (CPU tests): out = [cta_off (ncta + 1) | items {tile, first stripe, last stripe + 1} | tile_off (ntiles + 1) | slots],
 * off4 = the four offsets into out; returns the number of ints written, -needed when cap is too small, 0 for bad arguments. */
int sgp_debug_p2_plan(int ncta, int ntiles, int TM, long long slab_units, int w_diag, int w_off, int w_fixed, int* out, int cap, int* off4);
/* device pointers of the resident statistics of the last sweep: [psi2 (M*M) | psi1 (M*D_out) | psi0 | sum_y2] */
int sgp_stats_dev(sgp_ctx* ctx, double** psi2_dev, double** psi1_dev, double** scal_dev);

#ifdef __cplusplus
}
#endif
#endif /* SGP_H */
