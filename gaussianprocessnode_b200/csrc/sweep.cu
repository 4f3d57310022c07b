// Fused K_uf-tile generator + FP64 DMMA SYRK: the data sweep of the sparse variational GP node.
//
// Replaces the reference's per-data-point schedule -- `kernelmatrix!(Psi1_trans, kernel(theta), Xu, [x_n])` followed by
// the rank-1 `mul!(meta.Psi2, k, k', w, 0)` and the M x M add inside `prod` (GPnode/UniSGPnode.jl:144-173, 62-73;
// cubature variant GPnode/MultiSGPnode.jl:15-24) -- by one pass that never materialises K_uf in HBM.
//
// One persistent CTA per SM (256 threads = 8 warps, 2 per scheduler).  The work is the lower triangle of Psi2 cut into
// TM x TM tiles times the N range cut into chunks of NB = 32 points; every (tile, chunk) pair has a cost weight
// (diagonal tiles generate half the rows and skip the MMA blocks above the diagonal) and each CTA takes one contiguous,
// equally heavy slice of the (tile, chunk) sequence -- at most a few "segments" (tile, chunk range) per CTA.
//
// Per segment, software-pipelined over chunks (one __syncthreads per chunk):
//     1. cp.async.bulk (TMA, 1-D, SASS UBLKCP) stages the raw x / y / w blocks, 4 stages deep, mbarrier-tracked
//     2. raw block -> scaled records  x~ = sqrt(s) (x - c)/ell,  a = -|x~|^2/2                (s = 2048/ln 2)
//     3. every thread owns one inducing row and generates K_uf values of chunk c+1 into the other half of a
//        double-buffered shared-memory tile:   k = exp(a_n + b_m + x~_n . z~_m)   (16 FP64-pipe instructions per value
//        with the table-driven exp); diagonal tiles also fold Psi1 += k (w y) here
//     4. ... INTERLEAVED, k-step by k-step, with the DMMA.8x8x4 (mma.sync m8n8k4 f64 -- the only FP64 MMA sm_100a has)
//        consumption of chunk c into a 64x32 register accumulator per warp.  DFMA and DMMA share one pipe
//        (tools/fp64_microbench.cu), so the only way to keep it full is to feed it both streams at once: the DMMAs
//        hide the dependent-issue latency of the generator chains and vice versa.
//   After its last chunk a segment's partial tile goes to the workspace; a second kernel adds the partials of each
//   tile in a fixed order (deterministic, no FP64 atomics), mirrors the triangle and finishes Psi1.
//
// Roofline: FP64 DMMA pipe (measured 37.0 TFLOP/s on this pool's B200; cuBLAS DGEMM 35.5).  The generator's 16
// instructions per value are paid from the same budget: (TI+TJ)*16 / (TI*TJ) = 25 % on top of the MMA work for a 128x128
// off-diagonal tile.
#include "sweep_kernel.cuh"

using namespace sgp_sweep;

int sgp_sweep_chunk() { return 32; }

// Runs the whole sweep on device-resident data; statistics land in ctx->stats_dev.
int sgp_sweep_launch(sgp_ctx* ctx, const double* X, const double* y, const double* yv, const double* w, int64_t N, int64_t Ncap,
                     bool time_main) {
    if (!ctx->have_kernel || !ctx->have_Z) SGP_FAIL(ctx, SGP_ERR_ARG, "sweep: set_kernel and set_inducing first");
    constexpr int NB = 32;
    const int M = ctx->M, D = ctx->D;
    const int dpad = D <= 4 ? 4 : D <= 8 ? 8 : 16;
    const int TM = (M > 192) ? 128 : 64;
    const int nblk = (M + TM - 1) / TM;
    const int ntiles = nblk * (nblk + 1) / 2;
    const long long chunks = (N + NB - 1) / NB;
    if (chunks * NB > Ncap) SGP_FAIL(ctx, SGP_ERR_ARG, "sweep: data buffers must be padded to a multiple of 32 points");
    // one persistent CTA per SM; the (tile, chunk) sequence is cut into equally heavy contiguous slices.  Cost weights
    // of one chunk (measured per-chunk clocks, see profiles/): off-diagonal 2*TM generated rows + TM^2 MMA, diagonal TM
    // rows + the MMA blocks on or below the diagonal
    int w_diag = 15, w_off = 24, w_fixed = 17;
    if (const char* e = std::getenv("SGP_SWEEP_WEIGHTS")) {
        int a_ = 0, b_ = 0, c_ = -1;
        int got = std::sscanf(e, "%d,%d,%d", &a_, &b_, &c_);
        if (got >= 2 && a_ > 0 && b_ > 0) { w_diag = a_; w_off = b_; if (got == 3 && c_ >= 0) w_fixed = c_; }
    }
    const long long total_cost = chunks * ((long long)nblk * w_diag + (long long)(ntiles - nblk) * w_off) + (long long)ntiles * w_fixed;
    const int ncta = (int)std::min<long long>(std::min(ctx->num_sms, 256), std::max<long long>(1, chunks * ntiles));   // <= block size: see the segment count in run_segment
    const int grid = ncta;
    const int nslots = ncta + ntiles;

    size_t need_work = (size_t)nslots * TM * TM + (size_t)nslots * TM + (size_t)nslots * 2 + 2;
    int rc = sgp_ensure(ctx, &ctx->work_dev, &ctx->work_cap, need_work); if (rc) return rc;
    size_t need_stats = (size_t)M * M + (size_t)M + 8;
    rc = sgp_ensure(ctx, &ctx->stats_dev, &ctx->stats_cap, need_stats); if (rc) return rc;
    ctx->Dout = 1;

    SweepParams p{};
    p.X = X; p.y = y; p.yv = yv; p.w = w; p.N = N; p.chunks = chunks; p.M = M; p.D = D; p.ntiles = ntiles; p.nblk = nblk; p.ncta = ncta;
    p.total_cost = total_cost; p.w_diag = w_diag; p.w_off = w_off; p.w_fixed = w_fixed; p.dbg = (nslots <= 8192) ? ctx->sweep_dbg_dev : nullptr;
    if (p.dbg) { SGP_CUDA(ctx, cudaMemsetAsync(p.dbg, 0xff, (size_t)nslots * 4 * sizeof(long long), ctx->stream)); ctx->sweep_dbg_slots = nslots; }
    const double sq = std::sqrt(SGP_EXP_SCALE);
    for (int d = 0; d < SGP_MAX_D; ++d) { p.inv_ell_s[d] = d < D ? sq / ctx->ell[d] : 0.0; p.center[d] = d < D ? ctx->center[d] : 0.0; }
    p.log_var_s = SGP_EXP_SCALE * std::log(ctx->variance); p.variance = ctx->variance;
    p.Z = ctx->Z_dev; p.exptab = ctx->exptab_dev;
    p.partial = ctx->work_dev; p.psi1_partial = p.partial + (size_t)nslots * TM * TM; p.scal_partial = p.psi1_partial + (size_t)nslots * TM;
    p.psi2 = ctx->stats_dev; p.psi1 = p.psi2 + (size_t)M * M; p.scal = p.psi1 + M;

    int launches = 0;
    if (time_main) SGP_CUDA(ctx, cudaEventRecord(ctx->ev[2], ctx->stream));
    if (ctx->kind == SGP_KERNEL_SE) rc = launch_se(ctx, p, w != nullptr, grid, dpad, TM);
    else if (ctx->kind == SGP_KERNEL_MATERN32) rc = launch_m32(ctx, p, w != nullptr, grid, dpad, TM);
    else rc = launch_m52(ctx, p, w != nullptr, grid, dpad, TM);
    if (rc) return rc;
    ++launches;
    if (time_main) SGP_CUDA(ctx, cudaEventRecord(ctx->ev[3], ctx->stream));
    SGP_CUDA(ctx, cudaGetLastError());
    ctx->last_launches = launches;
    ctx->have_stats = true;
    return SGP_OK;
}
