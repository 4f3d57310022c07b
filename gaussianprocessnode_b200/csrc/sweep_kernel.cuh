// Device code of the fused sweep (see sweep.cu for the design notes).  Included by sweep.cu (host logic) and by the three
// per-kernel-family translation units sweep_se.cu / sweep_m32.cu / sweep_m52.cu, which only instantiate the launchers:
// splitting the template instantiations over translation units keeps the build parallel.
#pragma once
#include "sgp_internal.cuh"
#include <cmath>
#include <algorithm>
#include <cstdlib>
#include <type_traits>
#include <utility>
#include <cooperative_groups.h>

namespace sgp_sweep {


constexpr int kStages = 4;       // TMA stages of raw points
constexpr int kRecBufs = 3;      // record buffers: written two chunks ahead, read by the generator and by the MMA (weights)

struct SweepParams {
    const double* X; const double* y; const double* w;   // device, point-major, padded to a chunk multiple
    const double* yv;                                    // Var[y_n] or nullptr (only enters sum_n w (y^2 + yv))
    const double* Z;                                     // [M][D] raw inducing inputs
    const double* exptab;
    double* partial;                                     // [nslots][TM*TM]   slot = cta + tile
    double* psi1_partial;                                // [nslots][TM]
    double* scal_partial;                                // [nslots][2]       sum w, sum w (y^2 + yv) of tile 0's segments
    double *psi2, *psi1, *scal;                          // results: M x M column-major (full symmetric), M, {Psi0, sum_y2, sum_w, N}
    long long* dbg;                                      // optional [nslots][4]: chunks, clocks, diag, cta
    long long N;
    long long chunks;
    long long total_cost;                                // chunks * sum of tile weights
    int M, D, ntiles, nblk, ncta;
    int w_diag, w_off;                                   // cost weights of one chunk of a diagonal / off-diagonal tile
    int w_fixed;                                         // cost of entering a tile (segment prologue), charged at each tile's start
    double inv_ell_s[SGP_MAX_D];                         // sqrt(s) / ell_d  (both operands carry sqrt(s))
    double center[SGP_MAX_D];
    double log_var_s;                                    // s * ln sigma^2
    double variance;
};

// ---- work partition (shared by the sweep and the reduce kernels, so that both see the same segments) -------------
__host__ __device__ inline long long cta_pos(long long total_cost, int ncta, int b) { return total_cost / ncta * b + total_cost % ncta * b / ncta; }
// chunk range [lo, hi) that the cost interval [p0, p1) covers inside a tile whose cost prefix is `pre`
__host__ __device__ inline void seg_range(long long p0, long long p1, long long pre, int wt, int wfix, long long chunks, long long& lo, long long& hi) {
    long long d0 = p0 - pre - wfix, d1 = p1 - pre - wfix;      // the first wfix units of a tile's span are its entry cost
    lo = d0 <= 0 ? 0 : (d0 + wt - 1) / wt;
    hi = d1 <= 0 ? 0 : (d1 + wt - 1) / wt;
    if (lo > chunks) lo = chunks;
    if (hi > chunks) hi = chunks;
}

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
// DMMA the compiler may schedule freely between the generator's instructions (no `volatile`: the accumulators carry the
// dependences)
__device__ __forceinline__ void dmma884_nv(double& c0, double& c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// Shared-memory plan (doubles unless noted):
//   Kt    [2][NB][LD]       generated K_uf tile, point-major, double-buffered; LD = 2*TM + 4 keeps the DMMA fragment
//                           loads (and the generator's 16-byte stores) bank-conflict-free
//   tab   [2048]            2^(j/2048)
//   zrec  [2*TM][DPAD+1]    scaled inducing rows of blocks I and J;  zbias [2*TM] their b_m
//   rec   [3][NB][REC]      scaled point records: x~[DPAD], a, w*y, w, pad (REC = 12 or 20: conflict-free A fragments)
//   stage [kStages]: X raw [NB*D] | y [NB] | w [NB]
//   mbarrier full[kStages]
template <int TM, int NB, int DPAD>
struct Smem {
    static constexpr int LD = 2 * TM + 4;
    static constexpr int REC = DPAD <= 8 ? 12 : 20;
    static constexpr int ZR = DPAD + 1;
    static constexpr int STAGE = NB * SGP_MAX_D + 2 * NB;     // doubles per stage (X sized for the largest D)
    static constexpr size_t kt = 0;
    static constexpr size_t tab = kt + (size_t)2 * NB * LD;
    static constexpr size_t rec = tab + SGP_EXP_TAB;
    static constexpr size_t stage = rec + (size_t)kRecBufs * NB * REC;
    static constexpr size_t bars = stage + (size_t)kStages * STAGE;
    static constexpr size_t zbias = bars + kStages;
    static constexpr size_t zrec = zbias + (size_t)2 * TM;
    static constexpr size_t total_doubles = zrec + (size_t)2 * TM * ZR;
    static constexpr size_t bytes = total_doubles * sizeof(double);
};

// compile-time loop: f(std::integral_constant<int, 0>{}), ..., f(std::integral_constant<int, N-1>{})
template <class F, int... Is>
__device__ __forceinline__ void static_for_impl(F&& f, std::integer_sequence<int, Is...>) { (f(std::integral_constant<int, Is>{}), ...); }
template <int N, class F>
__device__ __forceinline__ void static_for(F&& f) { static_for_impl(f, std::make_integer_sequence<int, N>{}); }

struct SmemPtrs {
    double *Kt, *tab, *zrec, *zbias, *rec, *stage;
    unsigned long long* full;
};

// One segment: tile (I, J), chunks [c_begin, c_begin + nchunks) of the N range.  `g` = chunks this CTA has already
// pushed through the pipeline (selects stage / record / tile buffers and the mbarrier parity).
template <int TM, int NB, int DPAD, int NT, int KIND, bool WEIGHTED, bool DIAG>
__device__ __forceinline__ void run_segment(const SweepParams& p, const SmemPtrs& sm, const int I, const int J, const long long c_begin,
                                            const int nchunks, const unsigned g, const int slot) {
    using S = Smem<TM, NB, DPAD>;
    constexpr int LD = S::LD, REC = S::REC, ZR = S::ZR;
    constexpr int NWARPS = NT / 32;
    constexpr int WR = NT / 128;                     // warp grid: WR x 4 warps over the TM x TM CTA tile
    constexpr int WM = TM / WR, WN = TM / 4;
    constexpr int MI = WM / 8, NJ = WN / 8;          // 8x8 DMMA blocks per warp tile
    constexpr int ROWS = DIAG ? TM : 2 * TM;         // K_uf rows this tile needs per point: the "panel" [I-block | J-block]
    constexpr int RB = ROWS / (8 * NWARPS);          // 8-row blocks of the panel one warp generates
    constexpr int KS = NB / 4;                       // k-steps (4 points) per chunk
    constexpr int KQ = (DPAD + 3) / 4;               // DMMA k-quarters of the dot product x~ . z~
    constexpr int NPB = 4 / RB;                      // 8-point blocks per generator unit
    constexpr int KSPAN = KS / RB;                   // k-steps one generator unit is spread over
    constexpr int NMAT = (KIND == SGP_KERNEL_SE) ? 0 : 3;   // extra stages of the Matern kernels: r^2 -> sqrt -> exponent
    constexpr int NST = KQ + 8 + NMAT;               // stages of a generator unit
    constexpr int E0 = KQ + 1 + NMAT;                // first stage of the exp
    static_assert(RB == 1 || RB == 2 || RB == 4, "generator mapping");
    static_assert(NB == 32, "generator schedule");

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int D = p.D;
    const unsigned stage_bytes = (unsigned)(NB * D * 8 + NB * 8 + (WEIGHTED ? NB * 8 : 0));
    long long t_start = 0;
    if (p.dbg) t_start = clock64();

    auto issue = [&](int c) {   // thread 0 only: stage chunk c_begin + c
        const int s = (g + c) % kStages;
        double* st = sm.stage + (size_t)s * S::STAGE;
        const long long n0 = (c_begin + c) * NB;
        mbar_expect_tx(&sm.full[s], stage_bytes);
        tma_load_1d(st, p.X + n0 * D, NB * D * 8, &sm.full[s]);
        tma_load_1d(st + NB * SGP_MAX_D, p.y + n0, NB * 8, &sm.full[s]);
        if (WEIGHTED) tma_load_1d(st + NB * SGP_MAX_D + NB, p.w + n0, NB * 8, &sm.full[s]);
    };
    // raw staged block -> scaled records, ONE warp per chunk (the warps take turns): a lane owns a point.  The other
    // warps go straight on; the late warp catches up because the scheduler's FP64 pipe, not issue, is the bottleneck.
    auto prep = [&](int c) {
        const unsigned gc = g + c;
        const int s = gc % kStages;
        mbar_wait(&sm.full[s], (gc / kStages) & 1u);
        const double* st = sm.stage + (size_t)s * S::STAGE;
        double* r = sm.rec + (size_t)(gc % kRecBufs) * NB * REC + lane * REC;
        double a0 = 0.0, a1 = 0.0;
#pragma unroll
        for (int d = 0; d < DPAD; d += 2) {
            double v0 = 0.0, v1 = 0.0;
            if (d < D) v0 = (st[lane * D + d] - p.center[d]) * p.inv_ell_s[d];
            if (d + 1 < D) v1 = (st[lane * D + d + 1] - p.center[d + 1]) * p.inv_ell_s[d + 1];
            *reinterpret_cast<double2*>(r + d) = make_double2(v0, v1);
            a0 = fma(v0, v0, a0);
            a1 = fma(v1, v1, a1);
        }
        const long long n = (c_begin + c) * NB + lane;
        const double wn = WEIGHTED ? st[NB * SGP_MAX_D + NB + lane] : 1.0;
        r[DPAD] = (n < p.N) ? -0.5 * (a0 + a1) : -1.0e300;      // padded points generate exact zeros
        r[DPAD + 1] = wn * st[NB * SGP_MAX_D + lane];
        r[DPAD + 2] = wn;
        r[DPAD + 3] = 0.0;
    };

    __syncthreads();   // the previous segment is completely done with zrec / zbias / the tile buffers
    // inducing rows: block I -> panel rows [0,TM), block J -> rows [TM, 2TM);  z~ = sqrt(s) (z - c)/ell,
    // b = s (ln sigma^2 - |(z-c)/ell|^2 / 2); rows >= M are padding that generates exact zeros
    for (int r = tid; r < ROWS; r += NT) {
        const int gm = (r < TM ? I * TM + r : J * TM + (r - TM));
        double a = 0.0;
#pragma unroll
        for (int d = 0; d < DPAD; ++d) {
            double v = 0.0;
            if (gm < p.M && d < D) v = (p.Z[(size_t)gm * D + d] - p.center[d]) * p.inv_ell_s[d];
            sm.zrec[r * ZR + d] = v;
            a = fma(v, v, a);
        }
        sm.zbias[r] = (gm < p.M) ? ((KIND == SGP_KERNEL_SE ? p.log_var_s : 0.0) - 0.5 * a) : -1.0e300;
    }
    if (tid == 0)
        for (int c = 0; c < kStages - 1 && c < nchunks; ++c) issue(c);
    if (warp == 0) prep(0);
    if (warp == 1 && nchunks > 1) prep(1);
    __syncthreads();

    // ---- generator: the dot products x~ . z~ are themselves DMMAs ----------------------------------------------------
    // A warp owns RB 8-row blocks of the panel.  One 8-point x 8-row tile of exponents is   C = a_n + b_m  (DADD),
    // C += X~ Z~'  (KQ DMMA.8x8x4: A = 8 points x 4 dims from the records, B = 4 dims x 8 rows held in registers for the
    // whole segment), then the 7 steps of exp_scaled (sgp_internal.cuh) on the two values a thread holds, which are two
    // adjacent rows of one point = one 16-byte store into the point-major tile.  Four tiles (8 independent chains) form a
    // unit that is advanced one stage at a time between the DMMAs of the SYRK.  Compared with one DFMA chain per value
    // this needs no broadcast loads of x~ (12 instead of ~300 shared-memory loads per warp and chunk) at the same FP64
    // pipe cost.   Row i of A / C is point  8*pb + perm(i),  perm = 0,2,1,3,4,6,5,7:  the eight lanes of a quarter-warp
    // then store to points two apart, which LD = 4 (mod 16) maps to disjoint banks.
    const int grow0 = warp * (8 * RB);
    const int prow = ((lane >> 2) & 4) | ((lane >> 3) & 1) | ((lane >> 1) & 2);     // perm(lane / 4)
    double zf[RB][KQ];
#pragma unroll
    for (int rb = 0; rb < RB; ++rb)
#pragma unroll
        for (int kk = 0; kk < KQ; ++kk) {
            const int d = kk * 4 + (lane & 3);
            zf[rb][kk] = (d < DPAD) ? sm.zrec[(grow0 + rb * 8 + (lane >> 2)) * ZR + d] : 0.0;
        }
    double psi1_acc[RB][2];
#pragma unroll
    for (int rb = 0; rb < RB; ++rb) psi1_acc[rb][0] = psi1_acc[rb][1] = 0.0;

    struct Unit { double t[8], q[8], u[NMAT ? 8 : 1], xa[NPB][KQ]; int n[8]; };
    auto gen_stage = [&](auto st_tag, Unit& u, const double* __restrict__ rn, double* __restrict__ Kn, const int ui) {
        constexpr int st = decltype(st_tag)::value;
        const double MAGIC = 6755399441055744.0;            // 1.5 * 2^52
        const double C1 = 3.384507717577858e-04, C2 = 5.72744624517204e-08, C3 = 6.461528672932365e-12;
        // tile q of the unit: point block ui*NPB + q/RB, row block q%RB; this thread's point in a point block: prow
        if constexpr (st == 0) {
#pragma unroll
            for (int pb = 0; pb < NPB; ++pb) {
                const double* r = rn + ((ui * NPB + pb) * 8 + prow) * REC;
#pragma unroll
                for (int kk = 0; kk < KQ; ++kk) u.xa[pb][kk] = r[kk * 4 + (lane & 3)];
                const double an = r[DPAD];
#pragma unroll
                for (int rb = 0; rb < RB; ++rb) {
                    const double2 b2 = *reinterpret_cast<const double2*>(sm.zbias + grow0 + rb * 8 + 2 * (lane & 3));
                    u.t[2 * (pb * RB + rb)] = an + b2.x;
                    u.t[2 * (pb * RB + rb) + 1] = an + b2.y;
                }
            }
        } else if constexpr (st <= KQ) {
#pragma unroll
            for (int q = 0; q < 4; ++q) dmma884_nv(u.t[2 * q], u.t[2 * q + 1], u.xa[q / RB][st - 1], zf[q % RB][st - 1]);
        } else if constexpr (NMAT > 0 && st == KQ + 1) {
            // Matern: the accumulated exponent is -(s/2) r^2 (r = |(x - z)/ell|);  u^2 = nu' r^2, nu' = 3 or 5
            const double f = (KIND == SGP_KERNEL_MATERN32 ? -6.0 : -10.0) / SGP_EXP_SCALE;
#pragma unroll
            for (int c = 0; c < 8; ++c) u.q[c] = fmax(u.t[c] * f, 0.0);
        } else if constexpr (NMAT > 0 && st == KQ + 2) {
#pragma unroll
            for (int c = 0; c < 8; ++c) u.u[c] = sqrt(u.q[c]);
        } else if constexpr (NMAT > 0 && st == KQ + 3) {
#pragma unroll
            for (int c = 0; c < 8; ++c) u.t[c] = fma(u.u[c], -SGP_EXP_SCALE, p.log_var_s);     // s (ln sigma^2 - u)
        } else if constexpr (st == E0) {
#pragma unroll
            for (int c = 0; c < 8; ++c) u.q[c] = u.t[c] + MAGIC;
        } else if constexpr (st == E0 + 1) {
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                // t < -2.0e6 (result below exp(-677): flushed to zero) <=> sign set and magnitude above: compare the high
                // word as an unsigned integer (ALU pipe); hi word of -2.0e6 = 0xC13E8480
                const bool tiny = (unsigned)__double2hiint(u.t[c]) > 0xC13E8480u;
                const int nn = __double2loint(u.q[c]);
                u.n[c] = tiny ? (int)0x80000000 : nn;
                u.q[c] = u.q[c] - MAGIC;
            }
        } else if constexpr (st == E0 + 2) {
#pragma unroll
            for (int c = 0; c < 8; ++c) u.t[c] = u.t[c] - u.q[c];          // r
        } else if constexpr (st == E0 + 3) {
#pragma unroll
            for (int c = 0; c < 8; ++c) u.q[c] = fma(u.t[c], C3, C2);
        } else if constexpr (st == E0 + 4) {
#pragma unroll
            for (int c = 0; c < 8; ++c) u.q[c] = fma(u.q[c], u.t[c], C1);
        } else if constexpr (st == E0 + 5) {
#pragma unroll
            for (int c = 0; c < 8; ++c) u.q[c] = u.q[c] * u.t[c];
        } else {
            double T[8], res[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) T[c] = sm.tab[u.n[c] & (SGP_EXP_TAB - 1)];
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                res[c] = fma(T[c], u.q[c], T[c]);
                const int hi = __double2hiint(res[c]) + ((u.n[c] >> 11) << 20);
                res[c] = __hiloint2double(hi, __double2loint(res[c]));
                if (u.n[c] == (int)0x80000000) res[c] = 0.0;
                if (KIND == SGP_KERNEL_MATERN32) res[c] = fma(res[c], u.u[c], res[c]);                               // (1 + u) e^-u
                if (KIND == SGP_KERNEL_MATERN52) res[c] *= fma(u.u[c], fma(u.u[c], 1.0 / 3.0, 1.0), 1.0);           // (1 + u + u^2/3) e^-u
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int pt = (ui * NPB + q / RB) * 8 + prow;
                *reinterpret_cast<double2*>(Kn + pt * LD + grow0 + (q % RB) * 8 + 2 * (lane & 3)) = make_double2(res[2 * q], res[2 * q + 1]);
                if (DIAG) {
                    const double wy = rn[pt * REC + DPAD + 1];
                    psi1_acc[q % RB][0] = fma(res[2 * q], wy, psi1_acc[q % RB][0]);
                    psi1_acc[q % RB][1] = fma(res[2 * q + 1], wy, psi1_acc[q % RB][1]);
                }
            }
        }
    };

    // MMA mapping.  Off-diagonal tile: WR x 4 warps, a WM x WN register tile each.  Diagonal tile: only the lower triangle
    // is needed (the reduce kernel mirrors it), i.e. 10 of the 16 SB x SB sub-blocks (SB = TM/4).  They are dealt out so
    // that every scheduler (warps w and w+4) gets exactly 2.5: slot 0 = one sub-block for every warp (the off-diagonal ones to
    // warps 0-3, the diagonal ones to warps 4-7), slot 1 = a QUARTER sub-block (SB x 8) for every warp, cut from (2,0) and (3,0).
    // 40 DMMAs per k-step and scheduler instead of 64, 20 per warp.
    constexpr int SB = TM / 4, SI = SB / 8;
    static_assert(!DIAG || NWARPS == 8, "diagonal sub-block schedule is written for 8 warps");
    const int wr = warp >> 2, wc = warp & 3;
    const int s0r = (0x32103321 >> (4 * warp)) & 0xf, s0c = (0x32102110 >> (4 * warp)) & 0xf;   // warps 7..0: rows 3,2,1,0,3,3,2,1
    // slot 1: the two left-over sub-blocks (2,0) and (3,0) cut into SI column blocks of 8, one per warp (TM = 128: every warp carries 16 + 4
    // DMMAs per k-step, so the two warps of a scheduler stay balanced -- one warp alone reaches only 76 % of the DMMA rate)
    const int s1r = warp < SI ? 2 : 3, s1c0 = (warp % SI) * 8;
    const bool has1 = DIAG && warp < 2 * SI;
    const int a_off = (DIAG ? s0r * SB : wr * WM) + (lane >> 2);              // + 8*i
    const int b_off = (DIAG ? s0c * SB : TM + wc * WN) + (lane >> 2);         // + 8*j
    const int a1_off = s1r * SB + (lane >> 2), b1_off = s1c0 + (lane >> 2);
    const int kq = lane & 3;
    constexpr int AI = DIAG ? 2 * SI : MI, AJ = DIAG ? SI : NJ;      // accumulator blocks (diagonal: slot s = rows [s*SI, (s+1)*SI))
    double acc[AI][AJ][2];
#pragma unroll
    for (int i = 0; i < AI; ++i)
#pragma unroll
        for (int j = 0; j < AJ; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    // prologue: tile of chunk 0 (nothing to overlap with)
    {
        double* K0 = sm.Kt + (size_t)(g & 1u) * NB * LD;
        const double* r0 = sm.rec + (size_t)(g % kRecBufs) * NB * REC;
#pragma unroll 1
        for (int ui = 0; ui < RB; ++ui) {
            Unit u;
            static_for<NST>([&](auto st) { gen_stage(st, u, r0, K0, ui); });
        }
    }
    __syncthreads();

    // one pipeline step: consume chunk c (DMMA.8x8x4) while generating chunk c+1 (GEN)
    auto body = [&](auto gen_tag, const int c) {
        constexpr bool GEN = decltype(gen_tag)::value;
        constexpr int FI = DIAG ? SI : MI, FJ = DIAG ? SI : NJ;   // fragment blocks of the (slot-0) tile
        constexpr int FJ1 = 1;                                   // slot 1 of a diagonal tile: one 8-column block
        constexpr int DPK = FI * FJ;                  // DMMAs per k-step (diagonal: of slot 0)
        constexpr int TOTAL = KSPAN * DPK;            // DMMAs one generator unit is interleaved with
        static_assert(TOTAL >= NST, "generator schedule");
        const unsigned gc = g + c;
        const double* Kc = sm.Kt + (size_t)(gc & 1u) * NB * LD;
        double* Kn = sm.Kt + (size_t)((gc + 1) & 1u) * NB * LD;
        const double* rc = sm.rec + (size_t)(gc % kRecBufs) * NB * REC;
        const double* rn = sm.rec + (size_t)((gc + 1) % kRecBufs) * NB * REC;
        if (tid == 0 && c + kStages - 1 < nchunks) issue(c + kStages - 1);
        if (c + 2 < nchunks && warp == (int)(gc % NWARPS)) prep(c + 2);

#pragma unroll 1
        for (int ui = 0; ui < RB; ++ui) {
            const int ks0 = ui * KSPAN;
            Unit u;
            double a[FI], b[FJ];
            static_for<TOTAL>([&](auto d_tag) {
                constexpr int d = decltype(d_tag)::value;
                constexpr int kk = d / DPK, dd = d % DPK, i = dd / FJ, j = dd % FJ;
                if constexpr (dd == 0) {
                    const double* row = Kc + ((ks0 + kk) * 4 + kq) * LD;
                    if (DIAG && kk > 0 && has1) {      // slot 1 of the previous k-step (a, b are free to be overwritten)
                        const double* prow_ = row - 4 * LD;
                        double a1[FI], b1[FJ1];
#pragma unroll
                        for (int ii = 0; ii < FI; ++ii) a1[ii] = prow_[a1_off + 8 * ii];
#pragma unroll
                        for (int jj = 0; jj < FJ1; ++jj) b1[jj] = prow_[b1_off + 8 * jj];
                        if (WEIGHTED) {
                            const double wn = rc[((ks0 + kk - 1) * 4 + kq) * REC + DPAD + 2];
#pragma unroll
                            for (int jj = 0; jj < FJ1; ++jj) b1[jj] *= wn;
                        }
#pragma unroll
                        for (int ii = 0; ii < FI; ++ii)
#pragma unroll
                            for (int jj = 0; jj < FJ1; ++jj) dmma884_nv(acc[(DIAG ? FI : 0) + ii][jj][0], acc[(DIAG ? FI : 0) + ii][jj][1], a1[ii], b1[jj]);
                    }
#pragma unroll
                    for (int ii = 0; ii < FI; ++ii) a[ii] = row[a_off + 8 * ii];
#pragma unroll
                    for (int jj = 0; jj < FJ; ++jj) b[jj] = row[b_off + 8 * jj];
                    if (WEIGHTED) {
                        const double wn = rc[((ks0 + kk) * 4 + kq) * REC + DPAD + 2];
#pragma unroll
                        for (int jj = 0; jj < FJ; ++jj) b[jj] *= wn;
                    }
                }
                // the stage (at most one: TOTAL >= NST) whose slot floor(st * TOTAL / NST) is this DMMA
                constexpr int st = (d * NST + TOTAL - 1) / TOTAL;
#ifndef SGP_DBG_NOGEN
                if constexpr (GEN && st < NST && (st * TOTAL) / NST == d) gen_stage(std::integral_constant<int, st>{}, u, rn, Kn, ui);
#endif
#ifndef SGP_DBG_NOMMA
                dmma884_nv(acc[i][j][0], acc[i][j][1], a[i], b[j]);
#else
                acc[i][j][0] += a[i]; acc[i][j][1] += b[j];      // timing experiment only: keeps the fragment loads alive
#endif
            });
            if (has1) {      // slot 1 of the unit's last k-step
                const double* row = Kc + ((ks0 + KSPAN - 1) * 4 + kq) * LD;
                double a1[FI], b1[FJ1];
#pragma unroll
                for (int ii = 0; ii < FI; ++ii) a1[ii] = row[a1_off + 8 * ii];
#pragma unroll
                for (int jj = 0; jj < FJ1; ++jj) b1[jj] = row[b1_off + 8 * jj];
                if (WEIGHTED) {
                    const double wn = rc[((ks0 + KSPAN - 1) * 4 + kq) * REC + DPAD + 2];
#pragma unroll
                    for (int jj = 0; jj < FJ1; ++jj) b1[jj] *= wn;
                }
#pragma unroll
                for (int ii = 0; ii < FI; ++ii)
#pragma unroll
                    for (int jj = 0; jj < FJ1; ++jj) dmma884_nv(acc[(DIAG ? FI : 0) + ii][jj][0], acc[(DIAG ? FI : 0) + ii][jj][1], a1[ii], b1[jj]);
            }
        }
#ifndef SGP_DBG_NOBAR
        __syncthreads();   // chunk c consumed by every warp, chunk c+1 generated, records of chunk c+2 complete
#endif
    };
    for (int c = 0; c + 1 < nchunks; ++c) body(std::true_type{}, c);
    body(std::false_type{}, nchunks - 1);

    // ---- epilogue: register tile -> workspace slot (row-major TM x TM) ------------------------------------------
    double* out = p.partial + (size_t)slot * (TM * TM);
    if (DIAG) {
#pragma unroll
        for (int sl = 0; sl < 2; ++sl) {
            if (sl == 1 && !has1) break;
            const int r0 = (sl ? s1r : s0r) * SB, c0 = sl ? s1c0 : s0c * SB;
#pragma unroll
            for (int i = 0; i < SI; ++i)
#pragma unroll
                for (int j = 0; j < (sl ? 1 : SI); ++j) {
                    const int rr = r0 + 8 * i + (lane >> 2), cc = c0 + 8 * j + 2 * (lane & 3);
                    *reinterpret_cast<double2*>(out + rr * TM + cc) = make_double2(acc[sl * SI + i][j][0], acc[sl * SI + i][j][1]);
                }
        }
    } else {
#pragma unroll
        for (int i = 0; i < AI; ++i)
#pragma unroll
            for (int j = 0; j < AJ; ++j) {
                const int rr = wr * WM + 8 * i + (lane >> 2), cc = wc * WN + 8 * j + 2 * (lane & 3);
                *reinterpret_cast<double2*>(out + rr * TM + cc) = make_double2(acc[i][j][0], acc[i][j][1]);
            }
    }
    if (DIAG) {   // Psi1 rows of this warp: sum over the eight point positions (lane / 4), lanes 0..3 hold two rows each
#pragma unroll
        for (int rb = 0; rb < RB; ++rb)
#pragma unroll
            for (int v = 0; v < 2; ++v) {
                double x = psi1_acc[rb][v];
                x += __shfl_xor_sync(0xffffffffu, x, 4);
                x += __shfl_xor_sync(0xffffffffu, x, 8);
                x += __shfl_xor_sync(0xffffffffu, x, 16);
                if (lane < 4) p.psi1_partial[(size_t)slot * TM + grow0 + rb * 8 + 2 * lane + v] = x;
            }
    }
    if (DIAG && I == 0) {   // tile 0 sees every point once: sum_n w_n and sum_n w_n (y_n^2 + yv_n) of this segment's points
        const long long n_lo = c_begin * NB, n_hi = min(p.N, (c_begin + nchunks) * (long long)NB);
        double sw = 0.0, sy = 0.0;
        for (long long n = n_lo + tid; n < n_hi; n += NT) {
            const double wn = WEIGHTED ? p.w[n] : 1.0, yn = p.y[n], vn = p.yv ? p.yv[n] : 0.0;
            sw += wn;
            sy = fma(wn, fma(yn, yn, vn), sy);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            sw += __shfl_xor_sync(0xffffffffu, sw, o);
            sy += __shfl_xor_sync(0xffffffffu, sy, o);
        }
        double* red = sm.stage;            // the staging buffers are idle here
        __syncthreads();
        if (lane == 0) { red[2 * warp] = sw; red[2 * warp + 1] = sy; }
        __syncthreads();
        if (tid == 0) {
            double a = 0.0, b = 0.0;
            for (int wq = 0; wq < NWARPS; ++wq) { a += red[2 * wq]; b += red[2 * wq + 1]; }
            p.scal_partial[2 * slot] = a;
            p.scal_partial[2 * slot + 1] = b;
        }
    }
    if (p.dbg && tid == 0) {
        p.dbg[4 * slot + 0] = nchunks;
        p.dbg[4 * slot + 1] = clock64() - t_start;
        p.dbg[4 * slot + 2] = DIAG ? 1 : 0;
        p.dbg[4 * slot + 3] = blockIdx.x;
    }

}

// Second phase, after a grid-wide barrier: every CTA takes an equal share of the (tile, 4-row stripe) items, adds the
// partials of the tile's segments in slot (= CTA) order -- deterministic, no FP64 atomics --, mirrors the triangle into the
// full symmetric Psi2 and finishes Psi1 and the scalars.
template <int TM, int NT>
__device__ __forceinline__ void reduce_items(const SweepParams& p, double* __restrict__ S_, int* __restrict__ ibuf) {
    constexpr int SR = 4, STRIPES = TM / SR, NWARPS = NT / 32, LDS_ = TM + 1;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nitems = p.ntiles * STRIPES;
    const int it0 = (int)((long long)nitems * blockIdx.x / gridDim.x), it1 = (int)((long long)nitems * (blockIdx.x + 1) / gridDim.x);
    int* slots = ibuf + NWARPS;
    int cur_tile = -1, nseg = 0, I = 0, J = 0;
    for (int it = it0; it < it1; ++it) {
        const int tile = it / STRIPES, stripe = it - tile * STRIPES;
        if (tile != cur_tile) {      // the tile's workspace slots in CTA order: one candidate CTA per thread (ncta <= NT)
            cur_tile = tile;
            long long pre = 0;
            I = 0; J = 0;
            for (int t = 0; t < tile; ++t) {
                pre += (long long)(I == J ? p.w_diag : p.w_off) * p.chunks + p.w_fixed;
                if (++J > I) { ++I; J = 0; }
            }
            int mine = 0;
            if (tid < p.ncta) {
                long long lo, hi;
                seg_range(cta_pos(p.total_cost, p.ncta, tid), cta_pos(p.total_cost, p.ncta, tid + 1), pre, I == J ? p.w_diag : p.w_off, p.w_fixed, p.chunks, lo, hi);
                mine = lo < hi;
            }
            const unsigned ballot = __ballot_sync(0xffffffffu, mine);
            __syncthreads();                                     // the previous item is done with ibuf / S_
            if (lane == 0) ibuf[warp] = (int)ballot;
            nseg = __syncthreads_count(mine);
            if (mine) {
                int before = __popc(ballot & ((1u << lane) - 1u));
                for (int wq = 0; wq < warp; ++wq) before += __popc((unsigned)ibuf[wq]);
                slots[before] = tid + tile;
            }
        }
        __syncthreads();
        const bool diag = (I == J);
        const int r0 = stripe * SR;
        if (tid < SR * TM / 2) {
            const int e = 2 * tid, rl = e / TM, c = e - rl * TM, r = r0 + rl;
            double2 v = make_double2(0.0, 0.0);
            if (!diag || c <= r) {
                const double* src = p.partial + (size_t)r * TM + c;
#pragma unroll 8
                for (int sq = 0; sq < nseg; ++sq) {
                    const double2 x = __ldcg(reinterpret_cast<const double2*>(src + (size_t)slots[sq] * (TM * TM)));
                    v.x += x.x; v.y += x.y;
                }
            }
            S_[rl * LDS_ + c] = v.x; S_[rl * LDS_ + c + 1] = v.y;
        }
        __syncthreads();
        for (int e = tid; e < SR * TM; e += NT) {
            {   // psi2[gi + gj*M]: SR consecutive rows = one 32-byte sector per column
                const int rl = e % SR, c = e / SR, r = r0 + rl, gi = I * TM + r, gj = J * TM + c;
                if (gi < p.M && gj < p.M && (!diag || c <= r)) p.psi2[(size_t)gi + (size_t)gj * p.M] = S_[rl * LDS_ + c];
            }
            {   // mirror psi2[gj + gi*M]: consecutive threads -> consecutive columns
                const int rl = e / TM, c = e % TM, r = r0 + rl, gi = I * TM + r, gj = J * TM + c;
                if (gi < p.M && gj < p.M && (diag ? c < r : true)) p.psi2[(size_t)gj + (size_t)gi * p.M] = S_[rl * LDS_ + c];
            }
        }
        if (diag && tid < SR) {
            const int r = r0 + tid, gi = I * TM + r;
            if (gi < p.M) {
                double v = 0.0;
                for (int sq = 0; sq < nseg; ++sq) v += __ldcg(p.psi1_partial + (size_t)slots[sq] * TM + r);
                p.psi1[gi] = v;
            }
        }
        if (tile == 0 && stripe == 0 && tid == 0) {
            double sw = 0.0, sy = 0.0;
            for (int sq = 0; sq < nseg; ++sq) { sw += __ldcg(p.scal_partial + 2 * slots[sq]); sy += __ldcg(p.scal_partial + 2 * slots[sq] + 1); }
            p.scal[0] = p.variance * sw;   // Psi0 = sum_n w_n k(x_n, x_n)
            p.scal[1] = sy;                // sum_n w_n (ybar^2 + yvar)
            p.scal[2] = sw;
            p.scal[3] = (double)p.N;
        }
    }
}

template <int TM, int NB, int DPAD, int NT, int KIND, bool WEIGHTED>
__global__ void __launch_bounds__(NT, 1) sweep_kernel(const __grid_constant__ SweepParams p) {
    using S = Smem<TM, NB, DPAD>;
    extern __shared__ __align__(128) double smem[];
    SmemPtrs sm;
    sm.Kt = smem + S::kt; sm.tab = smem + S::tab; sm.zrec = smem + S::zrec; sm.rec = smem + S::rec; sm.stage = smem + S::stage;
    sm.zbias = smem + S::zbias;
    sm.full = reinterpret_cast<unsigned long long*>(smem + S::bars);

    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) mbar_init(&sm.full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    for (int i = tid; i < SGP_EXP_TAB; i += NT) sm.tab[i] = p.exptab[i];
    // (run_segment starts with a __syncthreads)

    const int bcta = blockIdx.x;
    const long long p0 = cta_pos(p.total_cost, p.ncta, bcta), p1 = cta_pos(p.total_cost, p.ncta, bcta + 1);
    unsigned g = 0;
    long long pre = 0;
    int I = 0, J = 0;
    for (int t = 0; t < p.ntiles && pre < p1; ++t) {
        const bool diag = (I == J);
        const int wt = diag ? p.w_diag : p.w_off;
        long long lo, hi;
        seg_range(p0, p1, pre, wt, p.w_fixed, p.chunks, lo, hi);
        if (lo < hi) {
            const int n = (int)(hi - lo);
            if (diag) run_segment<TM, NB, DPAD, NT, KIND, WEIGHTED, true>(p, sm, I, J, lo, n, g, bcta + t);
            else run_segment<TM, NB, DPAD, NT, KIND, WEIGHTED, false>(p, sm, I, J, lo, n, g, bcta + t);
            g += (unsigned)n;
        }
        pre += (long long)wt * p.chunks + p.w_fixed;
        if (++J > I) { ++I; J = 0; }
    }
    // every partial tile of the sweep is in the workspace once all CTAs are here (cooperative launch: all CTAs resident)
    __threadfence();
    cooperative_groups::this_grid().sync();
    reduce_items<TM, NT>(p, sm.Kt, reinterpret_cast<int*>(sm.stage));
}

template <int TM, int NB, int DPAD, int NT, int KIND>
int launch_t(sgp_ctx* ctx, const SweepParams& p, bool weighted, int grid) {
    using S = Smem<TM, NB, DPAD>;
    auto kern = weighted ? sweep_kernel<TM, NB, DPAD, NT, KIND, true> : sweep_kernel<TM, NB, DPAD, NT, KIND, false>;
    SGP_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S::bytes));
    void* args[] = {const_cast<SweepParams*>(&p)};
    SGP_CUDA(ctx, cudaLaunchCooperativeKernel((const void*)kern, dim3(grid), dim3(NT), args, S::bytes, ctx->stream));
    ctx->last_grid = grid; ctx->last_block = NT; ctx->last_smem = (int)S::bytes;
    SGP_CUDA(ctx, cudaGetLastError());
    return SGP_OK;
}

template <int TM, int NB, int NT, int KIND>
int launch_d(sgp_ctx* ctx, const SweepParams& p, bool weighted, int grid, int dpad) {
    switch (dpad) {
        case 4: return launch_t<TM, NB, 4, NT, KIND>(ctx, p, weighted, grid);
        case 8: return launch_t<TM, NB, 8, NT, KIND>(ctx, p, weighted, grid);
        default: return launch_t<TM, NB, 16, NT, KIND>(ctx, p, weighted, grid);
    }
}


template <int KIND>
int launch_kind(sgp_ctx* ctx, const SweepParams& p, bool weighted, int grid, int dpad, int TM) {
    constexpr int NB = 32, NT = 256;
    if (TM == 128) return launch_d<128, NB, NT, KIND>(ctx, p, weighted, grid, dpad);
    return launch_d<64, NB, NT, KIND>(ctx, p, weighted, grid, dpad);
}

// defined in sweep_se.cu / sweep_m32.cu / sweep_m52.cu
int launch_se(sgp_ctx* ctx, const SweepParams& p, bool weighted, int grid, int dpad, int TM);
int launch_m32(sgp_ctx* ctx, const SweepParams& p, bool weighted, int grid, int dpad, int TM);
int launch_m52(sgp_ctx* ctx, const SweepParams& p, bool weighted, int grid, int dpad, int TM);

}  // namespace sgp_sweep
