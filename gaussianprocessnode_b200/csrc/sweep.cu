// Fused K_uf-tile generator + FP64 DMMA SYRK: the data sweep of the sparse variational GP node.
//
// Replaces the reference's per-data-point schedule -- `kernelmatrix!(Psi1_trans, kernel(theta), Xu, [x_n])` followed by
// the rank-1 `mul!(meta.Psi2, k, k', w, 0)` and the M x M add inside `prod` (GPnode/UniSGPnode.jl:144-173, 62-73;
// cubature variant GPnode/MultiSGPnode.jl:15-24) -- by one pass that never materialises K_uf in HBM:
//
//   CTA (tile (I,J) of the lower triangle of Psi2, split s of the N range) loops over chunks of NB points:
//     1. cp.async.bulk (TMA, 1-D) stages the raw x / y / w blocks into shared memory, 4 stages deep, mbarrier-tracked
//     2. NB threads turn them into scaled records  x~ = s (x - c)/ell,  a = -s/2 |.|^2      (s = 2048/ln 2)
//     3. all 256 threads generate the K_uf tile rows of blocks I and J for the chunk straight into shared memory:
//        k = exp(a_n + b_m + x~_n . z~_m) (SE-ARD; 16 FP64 instructions per value with the table-driven exp) --
//        diagonal tiles also fold Psi1 += k * (w y) here
//     4. all 8 warps consume the tile with DMMA.8x8x4 (mma.sync m8n8k4 f64 -- the only FP64 MMA sm_100a has; the
//        m16n8k* PTX shapes are split into it by ptxas) into a 64x32 register accumulator per warp.
//   After the last chunk the 128x128 partial goes to the split-N workspace; a second kernel adds the splits in a fixed
//   order (deterministic, no FP64 atomics), mirrors the triangle and finishes Psi1.
//
// Roofline: FP64 DMMA pipe (measured 37.0 TFLOP/s on this pool's B200; cuBLAS DGEMM 35.5).  DFMA and DMMA share that
// pipe (tools/fp64_microbench.cu: mixed streams add up to ~35 TFLOP/s), so the generator's 16 instructions per value are
// paid from the same budget: (TI+TJ)*16 / (TI*TJ) = 25 % on top of the MMA work for a 128x128 tile.
#include "sgp_internal.cuh"
#include <cmath>
#include <algorithm>

namespace {

constexpr int kThreads = 512;
constexpr int kStages = 4;

struct SweepParams {
    const double* X; const double* y; const double* w;   // device, point-major, padded to a chunk multiple
    const double* zt;                                    // [Mpad][DPAD] scaled + centred inducing inputs
    const double* zb;                                    // [Mpad]       s * (ln sigma^2 - |z~|^2 / 2)
    const double* exptab;
    double* partial;                                     // [nsplit][ntiles][TM*TM]
    double* psi1_partial;                                // [nsplit][Mpad]
    long long N;
    long long chunks;
    int M, D, ntiles, nsplit, nblk;
    double inv_ell_s[SGP_MAX_D];                         // sqrt(s) / ell_d  (both operands carry sqrt(s))
    double center[SGP_MAX_D];
    double half_s_dummy;
};

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE;\n"
        "bra WAIT_LOOP;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// 1-D TMA bulk copy global -> shared, completion counted on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// Shared-memory plan (doubles unless noted):
//   Kt   [NB][LD]           generated K_uf tile, point-major; LD = 2*TM + 4 keeps DMMA fragment loads conflict-free
//   tab  [2048]             2^(j/2048)
//   zrec [2*TM][DPAD+1]     scaled inducing rows of blocks I and J (+ b_m)
//   rec  [2][NB][REC]       scaled point records (double-buffered): x~[DPAD], a, w*y, w, pad
//   stage[kStages]: X raw [NB*D] | y [NB] | w [NB]
//   mbarrier full[kStages]
template <int TM, int NB, int DPAD>
struct Smem {
    static constexpr int LD = 2 * TM + 4;
    static constexpr int REC = DPAD + 4;
    static constexpr int ZR = DPAD + 1;
    static constexpr int STAGE = NB * SGP_MAX_D + 2 * NB;     // doubles per stage (X sized for the largest D)
    static constexpr size_t kt = 0;
    static constexpr size_t tab = kt + (size_t)NB * LD;
    static constexpr size_t zrec = tab + SGP_EXP_TAB;
    static constexpr size_t rec = zrec + (size_t)2 * TM * ZR;
    static constexpr size_t stage = rec + (size_t)2 * NB * REC;
    static constexpr size_t bars = stage + (size_t)kStages * STAGE;
    static constexpr size_t red = bars + kStages;             // psi1 cross-group reduction [kThreads]
    static constexpr size_t total_doubles = red + kThreads;
    static constexpr size_t bytes = total_doubles * sizeof(double);
};

template <int TM, int NB, int DPAD, bool WEIGHTED>
__global__ void __launch_bounds__(kThreads, 1) sweep_kernel(const SweepParams p) {
    using S = Smem<TM, NB, DPAD>;
    constexpr int LD = S::LD, REC = S::REC, ZR = S::ZR;
    constexpr int WM = TM / 4, WN = TM / 4;          // warp tile: 4 x 4 warps over the TM x TM CTA tile
    constexpr int MI = WM / 8, NJ = WN / 8;          // 8x8 DMMA blocks per warp tile
    extern __shared__ __align__(128) double smem[];
    double* Kt = smem + S::kt;
    double* tab = smem + S::tab;
    double* zrec = smem + S::zrec;
    double* rec = smem + S::rec;                     // two record buffers: [2][NB][REC]
    double* stage = smem + S::stage;
    unsigned long long* full = reinterpret_cast<unsigned long long*>(smem + S::bars);
    double* red = smem + S::red;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tile = blockIdx.x % p.ntiles, split = blockIdx.x / p.ntiles;
    // lower-triangular tile index -> (I, J), I >= J
    int I = (int)((sqrtf(8.f * tile + 1.f) - 1.f) * 0.5f);
    while ((I + 1) * (I + 2) / 2 <= tile) ++I;
    while (I * (I + 1) / 2 > tile) --I;
    const int J = tile - I * (I + 1) / 2;
    const bool diag = (I == J);
    const int D = p.D;

    const long long c_begin = p.chunks * split / p.nsplit, c_end = p.chunks * (split + 1) / p.nsplit;
    const int nchunks = (int)(c_end - c_begin);
    const unsigned stage_bytes = (unsigned)(NB * D * 8 + NB * 8 + (WEIGHTED ? NB * 8 : 0));

    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    for (int i = tid; i < SGP_EXP_TAB; i += kThreads) tab[i] = p.exptab[i];
    // inducing rows: block I -> zrec rows [0,TM), block J -> rows [TM, 2TM)
    for (int i = tid; i < 2 * TM * ZR; i += kThreads) {
        int r = i / ZR, d = i - r * ZR;
        int g = (r < TM ? I * TM + r : J * TM + (r - TM));
        zrec[i] = (d < DPAD) ? p.zt[(size_t)g * DPAD + d] : p.zb[g];
    }
    __syncthreads();

    auto issue = [&](int c) {   // thread 0 only: stage chunk c_begin + c
        int s = c % kStages;
        double* st = stage + (size_t)s * S::STAGE;
        long long n0 = (c_begin + c) * NB;
        mbar_expect_tx(&full[s], stage_bytes);
        tma_load_1d(st, p.X + n0 * D, NB * D * 8, &full[s]);
        tma_load_1d(st + NB * SGP_MAX_D, p.y + n0, NB * 8, &full[s]);
        if (WEIGHTED) tma_load_1d(st + NB * SGP_MAX_D + NB, p.w + n0, NB * 8, &full[s]);
    };
    // raw staged block -> scaled records, spread over all threads: one (point, dimension) element each, |x~|^2 by a
    // shuffle reduction over the DPAD lanes of a point
    auto prep = [&](int c) {
        const int s = c % kStages;
        mbar_wait(&full[s], (unsigned)((c / kStages) & 1));
        const double* st = stage + (size_t)s * S::STAGE;
        double* rb = rec + (size_t)(c & 1) * NB * REC;
        for (int e = tid; e < NB * DPAD; e += kThreads) {      // NB*DPAD is a multiple of 32: whole warps take part
            const int pt = e / DPAD, d = e % DPAD;
            double v = 0.0;
            if (d < D) v = (st[pt * D + d] - p.center[d]) * p.inv_ell_s[d];
            rb[pt * REC + d] = v;
            double a = v * v;
#pragma unroll
            for (int o = DPAD / 2; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
            if (d == 0) {
                const long long n = (c_begin + c) * NB + pt;
                const double wn = WEIGHTED ? st[NB * SGP_MAX_D + NB + pt] : 1.0;
                rb[pt * REC + DPAD] = (n < p.N) ? -0.5 * a : -1.0e300;      // padded points generate exact zeros
                rb[pt * REC + DPAD + 1] = wn * st[NB * SGP_MAX_D + pt];
                rb[pt * REC + DPAD + 2] = wn;
            }
        }
    };
    if (tid == 0)
        for (int c = 0; c < kStages - 1 && c < nchunks; ++c) issue(c);

    // generator mapping: rows_needed rows (TM on the diagonal, 2*TM otherwise), tpr threads per row
    const int rows_needed = diag ? TM : 2 * TM;
    const int tpr = kThreads / rows_needed;
    const int grow = tid % rows_needed;                // row inside [I-block | J-block]
    const int ggrp = tid / rows_needed;
    const int npts = NB / tpr;                         // points per thread per chunk (multiple of 4)
    const int pt0 = ggrp * npts;
    double psi1_acc = 0.0;

    // MMA mapping
    const int wr = warp >> 2, wc = warp & 3;
    const int a_off = wr * WM + (lane >> 2);                        // + 8*i
    const int b_off = (diag ? 0 : TM) + wc * WN + (lane >> 2);      // + 8*j
    const int kq = lane & 3;
    double acc[MI][NJ][2];
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < NJ; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    if (nchunks > 0) prep(0);
    __syncthreads();

#ifdef SGP_PHASE_CLOCKS
    long long pc[6] = {0, 0, 0, 0, 0, 0}; long long t0c = clock64(), t1c;
#define PCLK(i) do { t1c = clock64(); pc[i] += t1c - t0c; t0c = t1c; } while (0)
#else
#define PCLK(i) do { } while (0)
#endif
    for (int c = 0; c < nchunks; ++c) {
        const double* rb = rec + (size_t)(c & 1) * NB * REC;
        // ---- generate: K_uf tile rows for blocks I and J ------------------------------------------------------
        {
            double zr[DPAD];
#pragma unroll
            for (int d = 0; d < DPAD; ++d) zr[d] = zrec[grow * ZR + d];
            const double zb = zrec[grow * ZR + DPAD];
            // U independent dependency chains per thread: the FP64 pipe has ~30 cycles of dependent-issue latency, so
            // the 16-instruction chain of one value is interleaved with others by hand (4 per thread x 4 warps per
            // scheduler = 16 chains in flight)
            constexpr int U = 4;
#pragma unroll 1
            for (int q = 0; q < npts; q += U) {
                const double* r = rb + (pt0 + q) * REC;
                double t[U];
#pragma unroll
                for (int u = 0; u < U; ++u) t[u] = r[u * REC + DPAD] + zb;
#pragma unroll
                for (int d = 0; d < DPAD; ++d)
#pragma unroll
                    for (int u = 0; u < U; ++u) t[u] = fma(r[u * REC + d], zr[d], t[u]);
                double k[U];
                exp_scaled_v<U>(t, k, tab);
#pragma unroll
                for (int u = 0; u < U; ++u) Kt[(pt0 + q + u) * LD + grow] = k[u];
                if (diag) {
#pragma unroll
                    for (int u = 0; u < U; ++u) psi1_acc = fma(k[u], r[u * REC + DPAD + 1], psi1_acc);
                }
            }
        }
        PCLK(0);
        __syncthreads();   // tile complete
        PCLK(1);
        // ---- stage + prepare the next chunk's records while the tile is consumed -------------------------------
        if (tid == 0 && c + kStages - 1 < nchunks) issue(c + kStages - 1);
        if (c + 1 < nchunks) prep(c + 1);
        PCLK(2);
        // ---- consume: Psi2 tile += K_I diag(w) K_J'  (DMMA.8x8x4) ----------------------------------------------
#pragma unroll 2
        for (int ks = 0; ks < NB / 4; ++ks) {
            const double* row = Kt + (ks * 4 + kq) * LD;
            double a[MI], b[NJ];
#pragma unroll
            for (int i = 0; i < MI; ++i) a[i] = row[a_off + 8 * i];
#pragma unroll
            for (int j = 0; j < NJ; ++j) b[j] = row[b_off + 8 * j];
            if (WEIGHTED) {
                const double wn = rb[(ks * 4 + kq) * REC + DPAD + 2];
#pragma unroll
                for (int j = 0; j < NJ; ++j) b[j] *= wn;
            }
#pragma unroll
            for (int i = 0; i < MI; ++i)
#pragma unroll
                for (int j = 0; j < NJ; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
        PCLK(3);
        __syncthreads();   // every warp is done with the tile and the next records are complete
        PCLK(4);
    }
#ifdef SGP_PHASE_CLOCKS
    if (lane == 0 && (warp == 0 || warp == 15) && (blockIdx.x == 0 || blockIdx.x == 77 || blockIdx.x == 8))
        printf("cta %d (tile %d diag %d) warp %d chunks %d: per-chunk clk gen %lld | wait B %lld | prep %lld | mma %lld | wait A %lld\n", blockIdx.x, tile,
               (int)diag, warp, nchunks, pc[0] / nchunks, pc[1] / nchunks, pc[2] / nchunks, pc[3] / nchunks, pc[4] / nchunks);
#endif

    // ---- epilogue: register tile -> split-N workspace (row-major TM x TM) ---------------------------------------
    double* out = p.partial + ((size_t)split * p.ntiles + tile) * (TM * TM);
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            int rr = wr * WM + 8 * i + (lane >> 2), cc = wc * WN + 8 * j + 2 * (lane & 3);
            *reinterpret_cast<double2*>(out + rr * TM + cc) = make_double2(acc[i][j][0], acc[i][j][1]);
        }
    if (diag) {
        red[tid] = psi1_acc;
        __syncthreads();
        if (tid < TM) {
            double v = 0.0;
            for (int g = 0; g < tpr; ++g) v += red[g * rows_needed + tid];
            p.psi1_partial[(size_t)split * (p.nblk * TM) + I * TM + tid] = v;
        }
    }
}

// Fixed-order sum over the N splits, mirror to the full symmetric matrix, finish Psi1.
template <int TM>
__global__ void reduce_kernel(const double* __restrict__ partial, const double* __restrict__ psi1_partial, double* __restrict__ psi2,
                              double* __restrict__ psi1, int M, int ntiles, int nsplit, int nblk) {
    const int tile = blockIdx.x;
    int I = (int)((sqrtf(8.f * tile + 1.f) - 1.f) * 0.5f);
    while ((I + 1) * (I + 2) / 2 <= tile) ++I;
    while (I * (I + 1) / 2 > tile) --I;
    const int J = tile - I * (I + 1) / 2;
    for (int e = threadIdx.x; e < TM * TM; e += blockDim.x) {
        int r = e / TM, c = e - r * TM;
        int gi = I * TM + r, gj = J * TM + c;
        if (gi >= M || gj >= M) continue;
        if (I == J && c > r) continue;
        double v = 0.0;
        for (int s = 0; s < nsplit; ++s) v += partial[((size_t)s * ntiles + tile) * (TM * TM) + e];
        psi2[(size_t)gi + (size_t)gj * M] = v;
        psi2[(size_t)gj + (size_t)gi * M] = v;
    }
    if (I == J) {
        for (int r = threadIdx.x; r < TM; r += blockDim.x) {
            int gi = I * TM + r;
            if (gi >= M) continue;
            double v = 0.0;
            for (int s = 0; s < nsplit; ++s) v += psi1_partial[(size_t)s * (nblk * TM) + gi];
            psi1[gi] = v;
        }
    }
}

// z~ = sqrt(s) (z - c)/ell,  b = s (ln sigma^2 - |(z-c)/ell|^2 / 2); rows >= M are padding that generates exact zeros.
__global__ void zprep_kernel(const double* __restrict__ Z, double* __restrict__ zt, double* __restrict__ zb, int M, int Mpad, int D,
                             int DPAD, SweepParams p, double log_var) {
    int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= Mpad) return;
    double a = 0.0;
    for (int d = 0; d < DPAD; ++d) {
        double v = 0.0;
        if (m < M && d < D) v = (Z[(size_t)m * D + d] - p.center[d]) * p.inv_ell_s[d];
        zt[(size_t)m * DPAD + d] = v;
        a = fma(v, v, a);
    }
    zb[m] = (m < M) ? (SGP_EXP_SCALE * log_var - 0.5 * a) : -1.0e300;
}

// sum_w, sum_w (y^2 + yv): warp-shuffle + block reduction, one partial per block, fixed-order finish
__global__ void scalar_sums_kernel(const double* __restrict__ y, const double* __restrict__ yv, const double* __restrict__ w, long long N,
                                   double* __restrict__ partial) {
    double sw = 0.0, sy = 0.0;
    for (long long n = blockIdx.x * (long long)blockDim.x + threadIdx.x; n < N; n += (long long)gridDim.x * blockDim.x) {
        double wn = w ? w[n] : 1.0, yn = y[n], vn = yv ? yv[n] : 0.0;
        sw += wn;
        sy = fma(wn, fma(yn, yn, vn), sy);
    }
    for (int o = 16; o > 0; o >>= 1) {
        sw += __shfl_xor_sync(0xffffffffu, sw, o);
        sy += __shfl_xor_sync(0xffffffffu, sy, o);
    }
    __shared__ double s_w[32], s_y[32];
    int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
    if (lane == 0) { s_w[wp] = sw; s_y[wp] = sy; }
    __syncthreads();
    if (wp == 0) {
        int nw = blockDim.x >> 5;
        sw = lane < nw ? s_w[lane] : 0.0;
        sy = lane < nw ? s_y[lane] : 0.0;
        for (int o = 16; o > 0; o >>= 1) {
            sw += __shfl_xor_sync(0xffffffffu, sw, o);
            sy += __shfl_xor_sync(0xffffffffu, sy, o);
        }
        if (lane == 0) { partial[2 * blockIdx.x] = sw; partial[2 * blockIdx.x + 1] = sy; }
    }
}
__global__ void scalar_finish_kernel(const double* __restrict__ partial, int nblocks, double variance, long long N, double* __restrict__ scal) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        double sw = 0.0, sy = 0.0;
        for (int b = 0; b < nblocks; ++b) { sw += partial[2 * b]; sy += partial[2 * b + 1]; }
        scal[0] = variance * sw;   // Psi0 = sum_n w_n k(x_n, x_n)
        scal[1] = sy;              // sum_n w_n (ybar^2 + yvar)
        scal[2] = sw;
        scal[3] = (double)N;
    }
}

template <int TM, int NB, int DPAD>
int launch_t(sgp_ctx* ctx, const SweepParams& p, bool weighted, int grid) {
    using S = Smem<TM, NB, DPAD>;
    auto kern = weighted ? sweep_kernel<TM, NB, DPAD, true> : sweep_kernel<TM, NB, DPAD, false>;
    SGP_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S::bytes));
    kern<<<grid, kThreads, S::bytes, ctx->stream>>>(p);
    ctx->last_grid = grid; ctx->last_block = kThreads; ctx->last_smem = (int)S::bytes;
    SGP_CUDA(ctx, cudaGetLastError());
    return SGP_OK;
}

template <int TM, int NB>
int launch_d(sgp_ctx* ctx, const SweepParams& p, bool weighted, int grid, int dpad) {
    switch (dpad) {
        case 2: return launch_t<TM, NB, 2>(ctx, p, weighted, grid);
        case 4: return launch_t<TM, NB, 4>(ctx, p, weighted, grid);
        case 8: return launch_t<TM, NB, 8>(ctx, p, weighted, grid);
        default: return launch_t<TM, NB, 16>(ctx, p, weighted, grid);
    }
}

}  // namespace

int sgp_sweep_chunk() { return 32; }

// Runs the whole sweep on device-resident data; statistics land in ctx->stats_dev.
int sgp_sweep_launch(sgp_ctx* ctx, const double* X, const double* y, const double* yv, const double* w, int64_t N, int64_t Ncap,
                     bool time_main) {
    if (!ctx->have_kernel || !ctx->have_Z) SGP_FAIL(ctx, SGP_ERR_ARG, "sweep: set_kernel and set_inducing first");
    if (ctx->kind != SGP_KERNEL_SE) SGP_FAIL(ctx, SGP_ERR_UNSUPPORTED, "fused sweep: SE-ARD only in this build (Matern: next)");
    constexpr int NB = 32;
    const int M = ctx->M, D = ctx->D;
    const int dpad = D <= 2 ? 2 : D <= 4 ? 4 : D <= 8 ? 8 : 16;
    const int TM = (M > 192) ? 128 : 64;
    const int nblk = (M + TM - 1) / TM, Mpad = nblk * TM;
    const int ntiles = nblk * (nblk + 1) / 2;
    const long long chunks = (N + NB - 1) / NB;
    if (chunks * NB > Ncap) SGP_FAIL(ctx, SGP_ERR_ARG, "sweep: data buffers must be padded to a multiple of 32 points");
    // split N so that tiles * splits fills the SMs: one wave of CTAs unless a CTA would still get >= 256 chunks in a
    // multi-wave launch (every CTA pays a fixed set-up -- exp table, inducing rows -- and a 128 KB partial tile)
    int best = 1; double best_eff = 0.0;
    for (int s = 1; s <= 64 && s <= chunks; ++s) {
        long long ctas = (long long)ntiles * s;
        long long waves = (ctas + ctx->num_sms - 1) / ctx->num_sms;
        if (waves > 1 && (s > 1 && chunks / s < 256)) break;
        if (waves > 4) break;
        double eff = (double)ctas / ((double)ctx->num_sms * (double)waves);
        if (eff > best_eff + 1e-9) { best_eff = eff; best = s; }
    }
    const int nsplit = best;
    const int grid = ntiles * nsplit;

    size_t need_work = (size_t)nsplit * ntiles * TM * TM + (size_t)nsplit * Mpad + 2 * 1024;
    int rc = sgp_ensure(ctx, &ctx->work_dev, &ctx->work_cap, need_work); if (rc) return rc;
    rc = sgp_ensure(ctx, &ctx->zrec_dev, &ctx->zrec_cap, (size_t)Mpad * (16 + 1)); if (rc) return rc;
    size_t need_stats = (size_t)M * M + (size_t)M + 8;
    rc = sgp_ensure(ctx, &ctx->stats_dev, &ctx->stats_cap, need_stats); if (rc) return rc;
    ctx->Dout = 1;

    SweepParams p{};
    p.X = X; p.y = y; p.w = w; p.N = N; p.chunks = chunks; p.M = M; p.D = D; p.ntiles = ntiles; p.nsplit = nsplit; p.nblk = nblk;
    const double sq = std::sqrt(SGP_EXP_SCALE);
    for (int d = 0; d < SGP_MAX_D; ++d) { p.inv_ell_s[d] = d < D ? sq / ctx->ell[d] : 0.0; p.center[d] = d < D ? ctx->center[d] : 0.0; }
    double* zt = ctx->zrec_dev; double* zb = ctx->zrec_dev + (size_t)Mpad * 16;
    p.zt = zt; p.zb = zb; p.exptab = ctx->exptab_dev;
    p.partial = ctx->work_dev; p.psi1_partial = ctx->work_dev + (size_t)nsplit * ntiles * TM * TM;
    double* scal_partial = p.psi1_partial + (size_t)nsplit * Mpad;
    double* psi2 = ctx->stats_dev; double* psi1 = psi2 + (size_t)M * M; double* scal = psi1 + M;

    int launches = 0;
    zprep_kernel<<<(Mpad + 127) / 128, 128, 0, ctx->stream>>>(ctx->Z_dev, zt, zb, M, Mpad, D, dpad, p, std::log(ctx->variance)); ++launches;
    if (time_main) SGP_CUDA(ctx, cudaEventRecord(ctx->ev[2], ctx->stream));
    if (TM == 128) rc = launch_d<128, NB>(ctx, p, w != nullptr, grid, dpad);
    else rc = launch_d<64, NB>(ctx, p, w != nullptr, grid, dpad);
    if (rc) return rc;
    ++launches;
    if (time_main) SGP_CUDA(ctx, cudaEventRecord(ctx->ev[3], ctx->stream));
    if (TM == 128) reduce_kernel<128><<<ntiles, 256, 0, ctx->stream>>>(p.partial, p.psi1_partial, psi2, psi1, M, ntiles, nsplit, nblk);
    else reduce_kernel<64><<<ntiles, 256, 0, ctx->stream>>>(p.partial, p.psi1_partial, psi2, psi1, M, ntiles, nsplit, nblk);
    ++launches;
    int sb = (int)std::min<long long>(1024, (N + 255) / 256); if (sb < 1) sb = 1;
    scalar_sums_kernel<<<sb, 256, 0, ctx->stream>>>(y, yv, w, N, scal_partial); ++launches;
    scalar_finish_kernel<<<1, 32, 0, ctx->stream>>>(scal_partial, sb, ctx->variance, N, scal); ++launches;
    SGP_CUDA(ctx, cudaGetLastError());
    ctx->last_launches = launches;
    ctx->have_stats = true;
    return SGP_OK;
}
