// The data sweep of the sparse variational GP node: host-side launch logic of the two sweep kernels.
//   sweep4_kernel.cuh  generate-once sweep (M > 384): K_uf generated once per sweep into an L2-resident ring of slab panels,
//                      consumed by a TMA-fed FP64 DMMA SYRK; see the notes at the top of that file and DESIGN.md section 4.1
//   sweep_kernel.cuh   the first fused kernel (K_uf tiles regenerated in shared memory inside every Psi2 tile): M <= 384, where few row
//                      blocks make the regeneration cheap, and the fallback for shapes outside the other one's range.  Its notes follow.
//
// Fused K_uf-tile generator + FP64 DMMA SYRK.
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
//     1. cp.async.bulk (TMA, 1-D, SASS UBLKCP) stages the raw x / y / w blocks in a multi-stage ring (kStages in sweep_kernel.cuh), mbarrier-tracked
//     2. raw block -> scaled records  x~ = sqrt(s) (x - c)/ell,  a = -|x~|^2/2                (s = 2048/ln 2)
//     3. every thread owns one inducing row and generates K_uf values of chunk c+1 into the other half of a
//        double-buffered shared-memory tile:   k = exp(a_n + b_m + x~_n . z~_m)   (16 FP64-pipe instructions per value
//        with the table-driven exp); diagonal tiles also fold Psi1 += k (w y) here
//     4. ... INTERLEAVED, k-step by k-step, with the DMMA.8x8x4 (mma.sync m8n8k4 f64 -- the only FP64 MMA sm_100a has)
//        consumption of chunk c into a 64x32 register accumulator per warp.  DFMA and DMMA share one pipe
//        (tools/fp64_microbench.cu), so the only way to keep it full is to feed it both streams at once: the DMMAs
//        hide the dependent-issue latency of the generator chains and vice versa.
//   After its last chunk a segment's partial tile goes to the workspace; after a grid barrier the same (cooperative) launch adds the
//   partials of each tile in a fixed order (deterministic, no FP64 atomics), mirrors the triangle and finishes Psi1.
//
// Roofline: FP64 DMMA pipe (measured 37.0 TFLOP/s on this pool's B200; cuBLAS DGEMM 35.5).  The generator's 16
// instructions per value are paid from the same budget: (TI+TJ)*16 / (TI*TJ) = 25 % on top of the MMA work for a 128x128
// off-diagonal tile.
#include "sweep4_kernel.cuh"
#include <cstring>
#include <vector>
#include <algorithm>

using namespace sgp_sweep;

int sgp_sweep_chunk() { return 32; }

// Phase-2 plan of the generate-once sweep, built once per configuration and kept on the device:
//   * per tile, the workspace slots (= cta + tile) of its segments in CTA order -- the same cta_pos / seg_range arithmetic the kernel uses for its
//     own segment table, evaluated here for all CTAs (the kernel used to redo it after the grid barrier: ~7 64-bit divisions per thread);
//   * per CTA, the (tile, stripe range) entries it reduces.  With at least as many CTAs as tiles every CTA gets stripes of ONE tile (a CTA whose
//     range straddled two tiles paid the slot set-up and a batch of loads twice and was the last to finish); the CTAs are dealt to the tiles in
//     proportion to the tiles' reduction cost (a stripe of a diagonal tile counts 1.25 off-diagonal ones -- half the partial loads, but its Psi1 rows ride along; measured, tools/p2_knob.py; SGP_SWEEP4_P2_DIAG).  More tiles than CTAs: contiguous
//     (tile, stripe) ranges, split at the tile boundaries.
// Every output element is still produced by one thread that adds the tile's partials in CTA order: bits do not depend on the plan.
// (pure host arithmetic: also exported as sgp_debug_p2_plan for the CPU tests)
static void p2_plan_host(int ncta, int ntiles, int TM, long long total_cost, long long slab_units, int w_diag, int w_off, int w_fixed, double diag_cost,
                         std::vector<int>& all, int off[4]) {
    const int SR = 4, STRIPES = TM / SR;
    std::vector<int> tile_off(ntiles + 1, 0), slots;
    {
        long long pre = 0;
        int I = 0, J = 0;
        for (int t = 0; t < ntiles; ++t) {
            const int wt = (I == J) ? w_diag : w_off;
            for (int c = 0; c < ncta; ++c) {
                long long lo, hi;
                seg_range(cta_pos(total_cost, ncta, c), cta_pos(total_cost, ncta, c + 1), pre, wt, w_fixed, slab_units, lo, hi);
                if (lo < hi) slots.push_back(c + t);
            }
            tile_off[t + 1] = (int)slots.size();
            pre += (long long)wt * slab_units + w_fixed;
            if (++J > I) { ++I; J = 0; }
        }
    }
    std::vector<int> cta_off(ncta + 1, 0), items;
    if (ncta >= ntiles) {
        // CTAs per tile in proportion to the reduction cost, at least one each; largest-remainder rounding keeps the total at ncta
        std::vector<double> want(ntiles);
        double tot = 0.0;
        { int I = 0, J = 0; for (int t = 0; t < ntiles; ++t) { want[t] = (I == J) ? diag_cost : 1.0; tot += want[t]; if (++J > I) { ++I; J = 0; } } }
        std::vector<int> cnt(ntiles, 1);
        int left = ncta - ntiles;
        for (int t = 0; t < ntiles; ++t) want[t] = want[t] / tot * ncta;
        for (int t = 0; t < ntiles; ++t) { const int extra = std::max(0, std::min(left, (int)want[t] - 1)); cnt[t] += extra; left -= extra; }
        while (left > 0) {      // the largest shortfall first (ties: the lower tile index -- deterministic)
            int best = 0;
            for (int t = 1; t < ntiles; ++t) if (want[t] - cnt[t] > want[best] - cnt[best]) best = t;
            ++cnt[best]; --left;
        }
        int c = 0;
        for (int t = 0; t < ntiles; ++t) {
            const int k = std::min(cnt[t], STRIPES);      // (more CTAs than stripes: the surplus idles in phase 2)
            for (int q = 0; q < cnt[t]; ++q, ++c) {
                cta_off[c] = (int)items.size() / 3;
                if (q < k) {
                    const int lo = STRIPES * q / k, hi = STRIPES * (q + 1) / k;
                    if (lo < hi) { items.push_back(t); items.push_back(lo); items.push_back(hi); }
                }
            }
        }
        cta_off[ncta] = (int)items.size() / 3;
    } else {
        const long long nitems = (long long)ntiles * STRIPES;
        for (int c = 0; c < ncta; ++c) {
            cta_off[c] = (int)items.size() / 3;
            long long it = nitems * c / ncta;
            const long long it1 = nitems * (c + 1) / ncta;
            while (it < it1) {
                const int tile = (int)(it / STRIPES), lo = (int)(it - (long long)tile * STRIPES);
                const int hi = (int)std::min<long long>(STRIPES, lo + (it1 - it));
                items.push_back(tile); items.push_back(lo); items.push_back(hi);
                it += hi - lo;
            }
        }
        cta_off[ncta] = (int)items.size() / 3;
    }
    all.clear();
    off[0] = 0; all.insert(all.end(), cta_off.begin(), cta_off.end());
    off[1] = (int)all.size(); all.insert(all.end(), items.begin(), items.end());
    off[2] = (int)all.size(); all.insert(all.end(), tile_off.begin(), tile_off.end());
    off[3] = (int)all.size(); all.insert(all.end(), slots.begin(), slots.end());
}

static double p2_diag_cost() {
    double diag_cost = 1.25;      // reduction cost of a diagonal tile's stripe relative to an off-diagonal one (half the partial loads, plus its Psi1 rows): measured
    if (const char* e = std::getenv("SGP_SWEEP4_P2_DIAG")) { const double v = std::atof(e); if (v > 0.05 && v < 20.0) diag_cost = v; }
    return diag_cost;
}

// The plan of a configuration as the kernel would get it, without a device (tests/test_host_helpers.py): out = [cta_off (ncta + 1) | items (3 per entry) |
// tile_off (ntiles + 1) | slots], off4 = the four offsets; returns the number of ints (or -needed when cap is too small).
extern "C" int sgp_debug_p2_plan(int ncta, int ntiles, int TM, long long slab_units, int w_diag, int w_off, int w_fixed, int* out, int cap, int* off4) {
    if (ncta < 1 || ntiles < 1 || (TM != 64 && TM != 128) || slab_units < 1 || !out || !off4) return 0;
    int nblk = 0;
    while ((nblk + 1) * (nblk + 2) / 2 <= ntiles) ++nblk;
    if (nblk * (nblk + 1) / 2 != ntiles) return 0;
    const long long total_cost = slab_units * ((long long)nblk * w_diag + (long long)(ntiles - nblk) * w_off) + (long long)ntiles * w_fixed;
    std::vector<int> all;
    p2_plan_host(ncta, ntiles, TM, total_cost, slab_units, w_diag, w_off, w_fixed, p2_diag_cost(), all, off4);
    if ((int)all.size() > cap) return -(int)all.size();
    std::memcpy(out, all.data(), all.size() * sizeof(int));
    return (int)all.size();
}

static int build_p2_plan(sgp_ctx* ctx, sgp_sweep4::Params& p, int TM) {
    const double diag_cost = p2_diag_cost();
    const long long key[8] = {p.total_cost, p.slab_units, p.ncta, p.ntiles, p.w_diag, p.w_off, ((long long)p.w_fixed << 20) + (long long)(diag_cost * 1000.0), TM};
    if (!ctx->p2plan_dev || std::memcmp(key, ctx->p2plan_key, sizeof(key)) != 0) {
        std::vector<int> all;
        int off[4];
        p2_plan_host(p.ncta, p.ntiles, TM, p.total_cost, p.slab_units, p.w_diag, p.w_off, p.w_fixed, diag_cost, all, off);
        if (ctx->p2plan_cap < all.size()) {
            if (ctx->p2plan_dev) SGP_CUDA(ctx, cudaFree(ctx->p2plan_dev));
            ctx->p2plan_dev = nullptr; ctx->p2plan_cap = 0;
            SGP_CUDA(ctx, cudaMalloc((void**)&ctx->p2plan_dev, (all.size() + 1024) * sizeof(int)));
            ctx->p2plan_cap = all.size() + 1024;
        }
        // (a pageable source: the runtime stages it before the call returns; an earlier sweep still reading the old plan is ordered before it on the stream)
        SGP_CUDA(ctx, cudaMemcpyAsync(ctx->p2plan_dev, all.data(), all.size() * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
        std::memcpy(ctx->p2plan_key, key, sizeof(key));
        std::memcpy(ctx->p2plan_off, off, sizeof(off));
    }
    p.p2_cta_off = ctx->p2plan_dev + ctx->p2plan_off[0]; p.p2_items = ctx->p2plan_dev + ctx->p2plan_off[1];
    p.p2_tile_off = ctx->p2plan_dev + ctx->p2plan_off[2]; p.p2_slots = ctx->p2plan_dev + ctx->p2plan_off[3];
    return SGP_OK;
}

// ---- generate-once sweep (sweep4_kernel.cuh): every K_uf value is generated once per sweep into an L2-resident slab panel and
// consumed by a TMA-fed DMMA SYRK.  Returns 1 when the shape is outside its range (nblk > CTAs): the caller falls back.
static int sweep_launch4(sgp_ctx* ctx, const double* X, const double* y, const double* yv, const double* w, int64_t N, bool time_main) {
    namespace s4 = sgp_sweep4;
    constexpr int NB = 32;
    const int M = ctx->M, D = ctx->D;
    const int dpad = D <= 4 ? 4 : D <= 8 ? 8 : 16;
    const int TM = (M > 192) ? 128 : 64;
    const int nblk = (M + TM - 1) / TM;
    const int ntiles = nblk * (nblk + 1) / 2;
    const long long chunks = (N + NB - 1) / NB;
    // slab: as many chunks as keep the panel (nblk x 32 x (TM + 4) doubles per chunk) inside the L2 budget
    double slab_mb = 20.0;        // per panel; the ring holds three: 60 MB is what stays in the L2 (persisting window) without DRAM re-reads
    if (const char* e = std::getenv("SGP_SWEEP_SLAB_MB")) { double v = std::atof(e); if (v > 0.0) slab_mb = v; }      // (forces the panel size)
    const size_t chunk_doubles = (size_t)nblk * NB * (TM + 4);
    long long max_slab = (long long)(slab_mb * 1048576.0 / (chunk_doubles * sizeof(double)));
    if (max_slab < 8) max_slab = 8;
    if (!std::getenv("SGP_SWEEP_SLAB_MB")) {
        // Small problems: when ALL of K_uf fits the L2 next to everything else (<= 64 MB: one slab; <= 96 MB: two), every panel is written once
        // and read once without any ring turnover -- one generator phase instead of several short ones (kin40k shape, 42 MB: 0.134 -> 0.119 ms;
        // N = 20 000: 0.246 -> 0.218 ms; tools/sweep_knobs.sh).  Larger problems keep the 3 x 20 MB ring that stays L2-resident.
        const double total_mb = (double)chunks * chunk_doubles * sizeof(double) / 1048576.0;
        if (total_mb <= 64.0) max_slab = chunks;
        else if (total_mb <= 96.0) max_slab = (chunks + 1) / 2;
    }
    const int nslabs = (int)((chunks + max_slab - 1) / max_slab);
    const long long slab_chunks = (chunks + nslabs - 1) / nslabs;
    int w_diag = 5, w_off = 8, w_fixed = 64;          // per k-step (4 points) / per segment, from the per-segment clock fit (tools/profile_sweep.py)
    if (const char* e = std::getenv("SGP_SWEEP4_WEIGHTS")) {
        int a_ = 0, b_ = 0, c_ = -1;
        int got = std::sscanf(e, "%d,%d,%d", &a_, &b_, &c_);
        if (got >= 2 && a_ > 0 && b_ > 0) { w_diag = a_; w_off = b_; if (got == 3 && c_ >= 0) w_fixed = c_; }
    }
    const long long slab_units = slab_chunks * (NB / 4);
    const long long total_cost = slab_units * ((long long)nblk * w_diag + (long long)(ntiles - nblk) * w_off) + (long long)ntiles * w_fixed;
    const int ncta = (int)std::min<long long>(std::min(ctx->num_sms, 256), std::max<long long>(1, slab_chunks * ntiles));
    if (nblk > ncta || ntiles / ncta + 2 > 60 || slab_units > 2000000000ll) return 1;
    const int nslots = ncta + ntiles;

    size_t need_work = (size_t)nslots * TM * TM + (size_t)ncta * TM + (size_t)ncta * 2 + 2;
    int rc = sgp_ensure(ctx, &ctx->work_dev, &ctx->work_cap, need_work); if (rc) return rc;
    const int nring = nslabs > 2 ? 3 : nslabs;
    rc = sgp_ensure(ctx, &ctx->kbuf_dev, &ctx->kbuf_cap, (size_t)nring * slab_chunks * chunk_doubles); if (rc) return rc;
    if (3 * (nblk + 1) > 1000) return 1;
    if (!ctx->sweep_flags_dev) {      // zeroed once; every sweep kernel leaves the counters zeroed when it ends
        SGP_CUDA(ctx, cudaMalloc((void**)&ctx->sweep_flags_dev, 1024 * sizeof(unsigned)));
        SGP_CUDA(ctx, cudaMemsetAsync(ctx->sweep_flags_dev, 0, 1024 * sizeof(unsigned), ctx->stream));
    }
    size_t need_stats = (size_t)M * M + (size_t)M + 8;
    rc = sgp_ensure_stats(ctx, need_stats); if (rc) return rc;
    ctx->Dout = 1;

    s4::Params p{};
    p.X = X; p.y = y; p.yv = yv; p.w = w; p.N = N; p.chunks = chunks; p.slab_chunks = slab_chunks; p.slab_units = slab_units; p.nslabs = nslabs;
    p.M = M; p.D = D; p.ntiles = ntiles; p.nblk = nblk; p.ncta = ncta;
    p.total_cost = total_cost; p.w_diag = w_diag; p.w_off = w_off; p.w_fixed = w_fixed; p.dbg = (nslots + 8 * ncta <= 8192) ? ctx->sweep_dbg_dev : nullptr;
    if (p.dbg) { SGP_CUDA(ctx, cudaMemsetAsync(p.dbg, 0xff, (size_t)(nslots + 8 * ncta) * 4 * sizeof(long long), ctx->stream)); ctx->sweep_dbg_slots = nslots + 8 * ncta; }
    const double sq = std::sqrt(SGP_EXP_SCALE);
    for (int d = 0; d < SGP_MAX_D; ++d) { p.inv_ell_s[d] = d < D ? sq / ctx->ell[d] : 0.0; p.center[d] = d < D ? ctx->center[d] : 0.0; }
    p.log_var_s = SGP_EXP_SCALE * std::log(ctx->variance); p.variance = ctx->variance;
    p.Z = ctx->Z_dev; p.exptab = ctx->exptab_dev; p.Kbuf = ctx->kbuf_dev; p.flags = ctx->sweep_flags_dev;
    p.slab_doubles = (long long)(slab_chunks * chunk_doubles); p.nring = nring;
    rc = build_p2_plan(ctx, p, TM); if (rc) return rc;
    // launch form: cooperative (co-residency of the persistent CTAs guaranteed by the driver) unless SGP_SWEEP_COOP=0 asks for a plain launch
    // with the kernel's own ticket barrier (the grid is one CTA per SM: resident as long as nothing else occupies the device)
    {
        const char* e = std::getenv("SGP_SWEEP_COOP");
        p.coop = (e && e[0] == '0') ? 0 : 1;
        p.gbar = ctx->sweep_flags_dev + 1000;          // (the dependency counters use the first 3 (nblk + 1) <= 1000 words)
        p.bar_epoch = ++ctx->sweep_bar_epoch;
    }
    p.partial = ctx->work_dev; p.psi1_partial = p.partial + (size_t)nslots * TM * TM; p.scal_partial = p.psi1_partial + (size_t)ncta * TM;
    p.psi2 = ctx->stats_dev; p.psi1 = p.psi2 + (size_t)M * M; p.scal = p.psi1 + M;
    // multi-GPU: phase 2 pushes this rank's statistics (Psi2 as the packed lower triangle) into its slot on every rank, and the kernel's tail adds
    // the slots in rank order once all ranks have published (xchg.cuh)
    ctx->last_sweep_exchanged = false;
    p.stats_out = ctx->stats_dev;
    if (ctx->want_exchange && sgp_comm_xchg(ctx, (size_t)M * (M + 1) / 2 + M + 4, &p.xr)) ctx->last_sweep_exchanged = true;
    // the host wants Psi2 as the packed lower triangle (sgp_sweep_psi_host_packed): the exchange leaves exactly that in this rank's result
    // buffer; without an exchange phase 2 writes a packed copy next to the full square
    p.packed_out = nullptr;
    if (ctx->want_packed) {
        if (ctx->last_sweep_exchanged) ctx->packed_src = reinterpret_cast<const double*>(p.xr.peers[p.xr.rank] + p.xr.slot0_off + p.xr.slot_bytes);
        else {
            rc = sgp_ensure(ctx, &ctx->packed_dev, &ctx->packed_cap, (size_t)M * (M + 1) / 2 + M + 8); if (rc) return rc;
            p.packed_out = ctx->packed_dev; ctx->packed_src = ctx->packed_dev;
        }
    }
    // an upload still in flight on the copy stream (sgp_sweep_psi_host): the kernel is launched NOW and waits on the device for the ready word that
    // follows the data -- launch latency and set-up run under the copies
    p.ready = nullptr; p.ready_val = 0;
    if (ctx->upload_pending) { p.ready = ctx->ready_dev; p.ready_val = ctx->ready_epoch; ctx->upload_pending = false; }

    // the panel ring is the only buffer worth keeping in L2: mark it persisting, everything else streams through the rest of the cache
    if (ctx->kbuf_window != (void*)ctx->kbuf_dev || ctx->kbuf_window_bytes != (size_t)nring * slab_chunks * chunk_doubles * sizeof(double)) {
        const size_t ring_bytes = (size_t)nring * slab_chunks * chunk_doubles * sizeof(double);
        int max_persist = 0, max_window = 0;
        cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, ctx->dev);
        cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, ctx->dev);
        const char* e = std::getenv("SGP_SWEEP_L2_PERSIST");
        if (max_persist > 0 && max_window > 0 && !(e && e[0] == '0')) {
            const size_t lim = std::min<size_t>((size_t)max_persist, ring_bytes);
            cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, lim);
            cudaStreamAttrValue av{};
            av.accessPolicyWindow.base_ptr = ctx->kbuf_dev;
            av.accessPolicyWindow.num_bytes = std::min<size_t>(ring_bytes, (size_t)max_window);
            av.accessPolicyWindow.hitRatio = (float)std::min(1.0, (double)lim / (double)av.accessPolicyWindow.num_bytes);
            av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
            av.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
            if (cudaStreamSetAttribute(ctx->stream, cudaStreamAttributeAccessPolicyWindow, &av) != cudaSuccess) cudaGetLastError();
        }
        ctx->kbuf_window = ctx->kbuf_dev; ctx->kbuf_window_bytes = ring_bytes;
    }

    if (time_main) SGP_CUDA(ctx, cudaEventRecord(ctx->ev[2], ctx->stream));
    if (ctx->kind == SGP_KERNEL_SE) rc = s4::launch4_se(ctx, p, w != nullptr, ncta, dpad, TM);
    else if (ctx->kind == SGP_KERNEL_MATERN32) rc = s4::launch4_m32(ctx, p, w != nullptr, ncta, dpad, TM);
    else rc = s4::launch4_m52(ctx, p, w != nullptr, ncta, dpad, TM);
    if (rc) return rc;
    if (time_main) SGP_CUDA(ctx, cudaEventRecord(ctx->ev[3], ctx->stream));
    SGP_CUDA(ctx, cudaGetLastError());
    ctx->last_launches = 1;
    ctx->have_stats = true;
    return SGP_OK;
}

// Runs the whole sweep on device-resident data; statistics land in ctx->stats_dev.
int sgp_sweep_launch(sgp_ctx* ctx, const double* X, const double* y, const double* yv, const double* w, int64_t N, int64_t Ncap,
                     bool time_main) {
    if (!ctx->have_kernel || !ctx->have_Z) SGP_FAIL(ctx, SGP_ERR_ARG, "sweep: set_kernel and set_inducing first");
    ctx->stats_of_data = false;      // set again by the caller that sweeps the resident data set (sgp_sweep_resident)
    ctx->packed_src = nullptr;       // EVERY sweep rewrites the resident statistics: a packed copy of the previous ones is stale (set again below on request)
    constexpr int NB = 32;
    {
        const long long chunks_ = (N + NB - 1) / NB;
        if (chunks_ * NB > Ncap) SGP_FAIL(ctx, SGP_ERR_ARG, "sweep: data buffers must be padded to a multiple of 32 points");
        // Which kernel: the generate-once sweep pays off when K_uf would otherwise be regenerated often -- nblk + 1 times with nblk row blocks
        // of 128 -- and when a slab carries enough tile work per CTA; measured crossover (tools/compare_sweep_kernels.py): nblk >= 4, i.e.
        // M > 384 (M = 512: 1.02-1.09x, 640: 1.07x, 1024: 1.2-1.3x; M <= 384: the first fused kernel is 8-20 % faster at large N).
        // SGP_SWEEP_IMPL=3 / 4 forces one of them.
        const char* impl = std::getenv("SGP_SWEEP_IMPL");
        const int nblk_ = (ctx->M + 127) / 128;
        const bool want4 = impl && impl[0] == '4' ? true : impl && impl[0] == '3' ? false : (ctx->M > 192 && nblk_ >= 4);
        if (want4) {
            int rc4 = sweep_launch4(ctx, X, y, yv, w, N, time_main);
            if (rc4 <= 0) return rc4;      // 1 = shape outside the generate-once kernel's range: fall through
        }
    }
    ctx->last_sweep_exchanged = false;
    ctx->packed_src = nullptr;
    { int rcj = sgp_join_upload(ctx); if (rcj) return rcj; }      // (this kernel has no ready-word wait)
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
    rc = sgp_ensure_stats(ctx, need_stats); if (rc) return rc;
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
