// Device code of the generate-once sweep (see sweep.cu for the design notes).
//
// The first fused kernel (sweep_kernel.cuh) regenerates the K_uf panel [I-block | J-block] inside every Psi2 tile that needs it:
// with nblk row blocks every K_uf value is generated nblk + 1 times, and because DFMA and DMMA share one FP64 pipe that
// regeneration is paid from the MMA budget (2 176 of 10 368 pipe clocks per off-diagonal chunk).  Here every K_uf value is
// generated exactly ONCE per sweep:
//
//   for each slab of the N range (sized so that its K_uf panel, M x slab points, stays in the 126 MB L2; a ring of three panels):
//     G phase   every CTA generates its share of the slab's panel -- (row block, chunk range) -- with the same DMMA-dot +
//               table-exp generator and writes it to the panel, laid out exactly as the consumer's shared-memory tile:
//               [chunk][row block][32 points][TM + 4]; Psi1 += k (w y) is folded in here.  Then it bumps the row block's
//               generation counter (release).
//     C phase   a plain TMA-fed DMMA SYRK: the segment waits (acquire) until the generation counters of its two row blocks
//               cover the slab, then per chunk ONE cp.async.bulk per row block (33 KB, already in the bank-conflict-free
//               layout) into a 3-stage mbarrier ring, fragments double-buffered in registers, 64x32 accumulators per warp
//               that are added into the CTA's workspace slot at the end of the segment.  Then the CTA bumps the consumption
//               counter; a panel of the ring is regenerated only when every CTA has consumed the slab that lived in it.
//   grid barrier, phase 2: deterministic reduction of the segment partials, mirror, Psi1, scalars.
//
// There is no grid-wide barrier inside the slab loop: CTAs run up to a slab apart, so load imbalance averages out over the
// sweep.  K_uf is never streamed from HBM: the ring's panels are rewritten in place slab after slab and are read from L2.
#pragma once
#include "sweep_kernel.cuh"
#include "xchg.cuh"

namespace sgp_sweep4 {

using sgp_sweep::cta_pos;
using sgp_sweep::seg_range;
using sgp_sweep::smem_u32;
using sgp_sweep::mbar_init;
using sgp_sweep::mbar_expect_tx;
using sgp_sweep::tma_load_1d;
using sgp_sweep::dmma884_nv;
using sgp_sweep::static_for;

constexpr int kKStages = 3;      // TMA stages of K_uf tiles (consumer)
constexpr int kXStages = 3;      // TMA stages of raw points (generator), one group of kGC chunks each
constexpr int kRecBufs = 2;      // record buffers of the generator (one group each)
constexpr int kGC = 4;           // chunks per generator group: one block barrier per group
constexpr int kMaxSeg = 64;      // segments of one CTA inside a slab (host falls back to the first kernel beyond)

struct Params {
    const double* X; const double* y; const double* w; const double* yv;
    const double* Z;
    const double* exptab;
    double* Kbuf;                 // [nring][slab_chunks][nblk][NB][TM + 4]   ring of K_uf panels (L2 resident)
    unsigned* flags;              // [nring][nblk generation counters | consumption counter]   (zeroed before the launch)
    long long slab_doubles;       // doubles per panel
    int nring;
    double* partial;              // [nslots][TM*TM]   slot = cta + tile
    double* psi1_partial;         // [ncta][TM]        Psi1 rows of the generator block of CTA i
    double* scal_partial;         // [ncta][2]         sum w, sum w (y^2 + yv) over the CTA's stripe of points
    double *psi2, *psi1, *scal;
    long long* dbg;
    long long N;
    long long chunks;             // all chunks of the sweep
    long long slab_chunks;        // chunks per slab
    long long slab_units;         // k-steps (4 points) per slab = 8 * slab_chunks: the work partition is defined on the k-steps of a full slab
    long long total_cost;
    int nslabs;
    int M, D, ntiles, nblk, ncta;
    int w_diag, w_off, w_fixed;
    double inv_ell_s[SGP_MAX_D];
    double center[SGP_MAX_D];
    double log_var_s;
    double variance;
    unsigned* gbar;               // {ticket, go} words of the hand-rolled grid barrier (plain launch: SGP_SWEEP_COOP=0)
    unsigned bar_epoch;           // ... value of `go` that releases this launch
    int coop;                     // 1: cooperative launch (grid.sync), 0: plain launch + ticket barrier
    SgpXchg xr;                   // multi-GPU exchange (nranks = 1: none): phase 2 then writes [packed lower triangle of Psi2 | Psi1 | scalars] into this rank's contribution buffer
    double* stats_out;            // ... and the sums over the ranks land here (full symmetric Psi2 | Psi1 | scalars); = psi2 without exchange
    double* packed_out;           // optional (no exchange): phase 2 also writes [packed lower triangle of Psi2 (LAPACK 'L' packed storage) | Psi1 | scalars] here, for the host
    // phase-2 plan, built on the host once per configuration (sweep.cu: build_p2_plan): which (tile, stripe range) every CTA reduces and, per tile,
    // the workspace slots of its segments in CTA order -- the kernel neither divides nor searches after the last grid barrier
    const int* p2_cta_off;        // [ncta + 1] offsets into p2_items
    const int* p2_items;          // {tile, first stripe, last stripe + 1} per entry
    const int* p2_tile_off;       // [ntiles + 1] offsets into p2_slots
    const int* p2_slots;          // workspace slots (= cta + tile) of the tile's segments, in CTA order
    const unsigned* ready;        // optional: a word the copy engine writes AFTER the input data of this sweep (sgp_sweep_psi_host): nothing reads
    unsigned ready_val;           // X / y / w / yv before *ready == ready_val
};

__device__ __forceinline__ void mbar_wait_(unsigned long long* bar, unsigned parity) {
    const unsigned a = smem_u32(bar);
    unsigned ok = 0;
    while (!ok) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n" : "=r"(ok) : "r"(a), "r"(parity) : "memory");
    }
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;\n" ::: "memory"); }
__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red_release_inc(unsigned* p) { asm volatile("red.release.gpu.global.add.u32 [%0], 1;\n" ::"l"(p) : "memory"); }
// one thread: wait until *flag >= target.  A wait that lasts seconds means a lost dependency: fail loudly instead of hanging the GPU.
__device__ __forceinline__ void spin_ge(const unsigned* flag, unsigned target) {
    if (ld_acquire(flag) >= target) return;
    const long long t0 = clock64();
    while (ld_acquire(flag) < target) {
        __nanosleep(32);
        if (clock64() - t0 > (1ll << 33)) __trap();
    }
}

// Shared memory (doubles): tab[2048] | barriers | union { consumer: stage[3] = { Ki [NB][LDB] | Kj [NB][LDB] | w [NB] }
//                                                        generator: xstage[3] (one group of 128 raw points each) | rec[2][128][REC] | zrec[TM][ZR] | zbias[TM]
//                                                        phase 2: S_ | ibuf }
template <int TM, int NB, int DPAD>
struct Smem4 {
    static constexpr int LDB = TM + 4;                         // = 4 (mod 16): conflict-free DMMA fragment loads
    static constexpr int REC = DPAD <= 8 ? 12 : 20;
    static constexpr int ZR = DPAD + 1;
    static constexpr int GP = kGC * NB;                        // points per generator group
    static constexpr int XSTAGE = GP * SGP_MAX_D + 2 * GP;
    static constexpr int KSTAGE = 2 * NB * LDB + NB;
    static constexpr size_t tab = 0;
    static constexpr size_t bars = tab + SGP_EXP_TAB;          // full[3] empty[3] xfull[3]
    static constexpr size_t segtab = bars + 16;                // this CTA's segments: kMaxSeg x {I, J, slot, lo, hi, -, -, -} (ints)
    static constexpr size_t u = segtab + kMaxSeg * 4;
    // consumer view
    static constexpr size_t kst = u;
    // generator view
    static constexpr size_t xstage = u;
    static constexpr size_t rec = xstage + (size_t)kXStages * XSTAGE;
    static constexpr size_t zrec = rec + (size_t)kRecBufs * GP * REC;
    static constexpr size_t zbias = zrec + (size_t)TM * ZR;
    static constexpr size_t gen_end = zbias + TM;
    static constexpr size_t con_end = kst + (size_t)kKStages * KSTAGE;
    static constexpr size_t total_doubles = con_end > gen_end ? con_end : gen_end;
    static constexpr size_t bytes = total_doubles * sizeof(double);
    static_assert((NB * LDB * 8) % 16 == 0 && (KSTAGE * 8) % 16 == 0 && (u * 8) % 128 == 0, "TMA alignment");
};

__host__ __device__ inline int gen_lo(int b, int ncta, int nblk);

struct Sm4 {
    double *tab, *u;
    int* segtab;
    unsigned long long *full, *empty, *xfull;
};

// thread 0: raw points of generator group gi (kGC chunks from chunk c_lo + gi * kGC; the last group may be shorter) -> X stage (xg + gi) % kXStages
template <int TM, int NB, int DPAD, bool WEIGHTED>
__device__ __forceinline__ void issue_points(const Params& p, const Sm4& sm, const long long c_lo, const int nchunks, const int gi, const unsigned xg) {
    using S = Smem4<TM, NB, DPAD>;
    constexpr int GP = S::GP;
    const int D = p.D;
    const int s = (xg + gi) % kXStages;
    double* st = sm.u + (S::xstage - S::u) + (size_t)s * S::XSTAGE;
    const long long n0 = (c_lo + (long long)gi * kGC) * NB;
    const unsigned npts = (unsigned)min(kGC, nchunks - gi * kGC) * NB;
    mbar_expect_tx(&sm.xfull[s], npts * (unsigned)(D * 8 + 8 + (WEIGHTED ? 8 : 0)));
    tma_load_1d(st, p.X + n0 * D, npts * D * 8, &sm.xfull[s]);
    tma_load_1d(st + GP * SGP_MAX_D, p.y + n0, npts * 8, &sm.xfull[s]);
    if (WEIGHTED) tma_load_1d(st + GP * SGP_MAX_D + GP, p.w + n0, npts * 8, &sm.xfull[s]);
}

// ---- G phase: this CTA's share of the slab's K_uf panel: row block `blk`, absolute chunks [c_lo, c_hi) -------------------------
template <int TM, int NB, int DPAD, int NT, int KIND, bool WEIGHTED>
__device__ __forceinline__ void gen_phase(const Params& p, const Sm4& sm, const int blk, const long long c_lo, const long long c_hi,
                                          const long long slab_c0, double* __restrict__ panel, unsigned& xg,
                                          double (&psi1_acc)[TM / (NT / 4)][2], const bool preissued) {
    using S = Smem4<TM, NB, DPAD>;
    constexpr int LDB = S::LDB, REC = S::REC, ZR = S::ZR;
    constexpr int NWARPS = NT / 32;
    constexpr int RB = TM / (8 * NWARPS);            // 8-row blocks of the row block one warp generates
    constexpr int KQ = (DPAD + 3) / 4;
    constexpr int NPB = 4 / RB;
    constexpr int NMAT = (KIND == SGP_KERNEL_SE) ? 0 : 3;
    constexpr int NST = KQ + 8 + NMAT;
    constexpr int E0 = KQ + 1 + NMAT;
    static_assert(RB == 1 || RB == 2 || RB == 4, "generator mapping");
    static_assert(NB == 32, "generator schedule");
    const int nchunks = (int)(c_hi - c_lo);
    if (nchunks <= 0) return;                        // CTA-uniform

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int D = p.D;
    double* xstage = sm.u + (S::xstage - S::u);
    double* rec = sm.u + (S::rec - S::u);
    double* zrec = sm.u + (S::zrec - S::u);
    double* zbias = sm.u + (S::zbias - S::u);
    constexpr int GP = S::GP;
    const int ngroups = (nchunks + kGC - 1) / kGC;

    auto issue = [&](int gi) { issue_points<TM, NB, DPAD, WEIGHTED>(p, sm, c_lo, nchunks, gi, xg); };   // thread 0 only
    // raw staged group -> scaled records, all warps: a half-warp lane owns a point, the two half-warps split the dimensions
    auto prep = [&](int gi) {
        const unsigned gg = xg + gi;
        const int s = gg % kXStages;
        mbar_wait_(&sm.xfull[s], (gg / kXStages) & 1u);
        const double* st = xstage + (size_t)s * S::XSTAGE;
        const int pt = warp * 16 + (lane & 15), half = lane >> 4;
        const int npts = min(kGC, nchunks - gi * kGC) * NB;
        double* r = rec + (size_t)(gg % kRecBufs) * GP * REC + pt * REC;
        double a0 = 0.0;
        if (pt < npts) {
#pragma unroll
            for (int dd = 0; dd < DPAD / 2; dd += 2) {
                const int d = half * (DPAD / 2) + dd;
                double v0 = 0.0, v1 = 0.0;
                if (d < D) v0 = (st[pt * D + d] - p.center[d]) * p.inv_ell_s[d];
                if (d + 1 < D) v1 = (st[pt * D + d + 1] - p.center[d + 1]) * p.inv_ell_s[d + 1];
                *reinterpret_cast<double2*>(r + d) = make_double2(v0, v1);
                a0 = fma(v0, v0, a0);
                a0 = fma(v1, v1, a0);
            }
        }
        a0 += __shfl_xor_sync(0xffffffffu, a0, 16);
        if (pt < npts && half == 0) {
            const long long n = (c_lo + (long long)gi * kGC) * NB + pt;
            const double wn = WEIGHTED ? st[GP * SGP_MAX_D + GP + pt] : 1.0;
            r[DPAD] = (n < p.N) ? -0.5 * a0 : -1.0e300;      // padded points generate exact zeros
            r[DPAD + 1] = wn * st[GP * SGP_MAX_D + pt];
            r[DPAD + 2] = wn;
            r[DPAD + 3] = 0.0;
        }
    };

    __syncthreads();   // the union region is free (previous phase complete in this CTA)
    for (int r = tid; r < TM; r += NT) {
        const int gm = blk * TM + r;
        double a = 0.0;
#pragma unroll
        for (int d = 0; d < DPAD; ++d) {
            double v = 0.0;
            if (gm < p.M && d < D) v = (p.Z[(size_t)gm * D + d] - p.center[d]) * p.inv_ell_s[d];
            zrec[r * ZR + d] = v;
            a = fma(v, v, a);
        }
        zbias[r] = (gm < p.M) ? ((KIND == SGP_KERNEL_SE ? p.log_var_s : 0.0) - 0.5 * a) : -1.0e300;
    }
    if (tid == 0 && !preissued)
        for (int gi = 0; gi < kXStages - 1 && gi < ngroups; ++gi) issue(gi);
    __syncwarp();
    prep(0);
    __syncthreads();

    const int grow0 = warp * (8 * RB);
    const int prow = ((lane >> 2) & 4) | ((lane >> 3) & 1) | ((lane >> 1) & 2);     // perm(lane / 4), see sweep_kernel.cuh
    double zf[RB][KQ];
#pragma unroll
    for (int rb = 0; rb < RB; ++rb)
#pragma unroll
        for (int kk = 0; kk < KQ; ++kk) {
            const int d = kk * 4 + (lane & 3);
            zf[rb][kk] = (d < DPAD) ? zrec[(grow0 + rb * 8 + (lane >> 2)) * ZR + d] : 0.0;
        }

    struct Unit { double t[8], q[8], u[NMAT ? 8 : 1], xa[NPB][KQ]; int n[8]; };
    auto gen_stage = [&](auto st_tag, Unit& u, const double* __restrict__ rn, double* __restrict__ Kg, const int ui) {
        constexpr int st = decltype(st_tag)::value;
        const double MAGIC = 6755399441055744.0;            // 1.5 * 2^52
        const double C1 = 3.384507717577858e-04, C2 = 5.72744624517204e-08, C3 = 6.461528672932365e-12;
        if constexpr (st == 0) {
#pragma unroll
            for (int pb = 0; pb < NPB; ++pb) {
                const double* r = rn + ((ui * NPB + pb) * 8 + prow) * REC;
#pragma unroll
                for (int kk = 0; kk < KQ; ++kk) u.xa[pb][kk] = r[kk * 4 + (lane & 3)];
                const double an = r[DPAD];
#pragma unroll
                for (int rb = 0; rb < RB; ++rb) {
                    const double2 b2 = *reinterpret_cast<const double2*>(zbias + grow0 + rb * 8 + 2 * (lane & 3));
                    u.t[2 * (pb * RB + rb)] = an + b2.x;
                    u.t[2 * (pb * RB + rb) + 1] = an + b2.y;
                }
            }
        } else if constexpr (st <= KQ) {
#pragma unroll
            for (int q = 0; q < 4; ++q) dmma884_nv(u.t[2 * q], u.t[2 * q + 1], u.xa[q / RB][st - 1], zf[q % RB][st - 1]);
        } else if constexpr (NMAT > 0 && st == KQ + 1) {
            const double f = (KIND == SGP_KERNEL_MATERN32 ? -6.0 : -10.0) / SGP_EXP_SCALE;
#pragma unroll
            for (int c = 0; c < 8; ++c) u.q[c] = fmax(u.t[c] * f, 0.0);
        } else if constexpr (NMAT > 0 && st == KQ + 2) {
#pragma unroll
            for (int c = 0; c < 8; ++c) u.u[c] = sqrt(u.q[c]);
        } else if constexpr (NMAT > 0 && st == KQ + 3) {
#pragma unroll
            for (int c = 0; c < 8; ++c) u.t[c] = fma(u.u[c], -SGP_EXP_SCALE, p.log_var_s);
        } else if constexpr (st == E0) {
#pragma unroll
            for (int c = 0; c < 8; ++c) u.q[c] = u.t[c] + MAGIC;
        } else if constexpr (st == E0 + 1) {
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const bool tiny = (unsigned)__double2hiint(u.t[c]) > 0xC13E8480u;      // t < -2.0e6: flushed to zero
                const int nn = __double2loint(u.q[c]);
                u.n[c] = tiny ? (int)0x80000000 : nn;
                u.q[c] = u.q[c] - MAGIC;
            }
        } else if constexpr (st == E0 + 2) {
#pragma unroll
            for (int c = 0; c < 8; ++c) u.t[c] = u.t[c] - u.q[c];
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
                if (KIND == SGP_KERNEL_MATERN32) res[c] = fma(res[c], u.u[c], res[c]);
                if (KIND == SGP_KERNEL_MATERN52) res[c] *= fma(u.u[c], fma(u.u[c], 1.0 / 3.0, 1.0), 1.0);
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int pt = (ui * NPB + q / RB) * 8 + prow;
#ifndef SGP4_DBG_NOSTORE
                *reinterpret_cast<double2*>(Kg + pt * LDB + grow0 + (q % RB) * 8 + 2 * (lane & 3)) = make_double2(res[2 * q], res[2 * q + 1]);
#endif
                const double wy = rn[pt * REC + DPAD + 1];
                psi1_acc[q % RB][0] = fma(res[2 * q], wy, psi1_acc[q % RB][0]);
                psi1_acc[q % RB][1] = fma(res[2 * q + 1], wy, psi1_acc[q % RB][1]);
            }
        }
    };

#pragma unroll 1
    for (int gi = 0; gi < ngroups; ++gi) {
        const unsigned gg = xg + gi;
        if (tid == 0 && gi + kXStages - 1 < ngroups) issue(gi + kXStages - 1);
        __syncwarp();
        if (gi + 1 < ngroups) prep(gi + 1);
        const int nc = min(kGC, nchunks - gi * kGC);
#pragma unroll 1
        for (int cc = 0; cc < nc; ++cc) {
            const double* rn = rec + (size_t)(gg % kRecBufs) * GP * REC + (size_t)cc * NB * REC;
            double* Kg = panel + ((size_t)(c_lo + (long long)gi * kGC + cc - slab_c0) * p.nblk + blk) * (size_t)(NB * LDB);
            if constexpr (RB == 1) {
                Unit u;
                static_for<NST>([&](auto st) { gen_stage(st, u, rn, Kg, 0); });
            } else {
                // two units in flight: their stages alternate, so that a warp always has 16 independent chains
                static_assert(RB == 2, "generator units");
                Unit u0, u1;
                static_for<NST>([&](auto st) { gen_stage(st, u0, rn, Kg, 0); gen_stage(st, u1, rn, Kg, 1); });
            }
        }
        __syncthreads();   // records of group gi+1 complete, group gi's records and the stage of group gi-1... free
    }
    xg += (unsigned)ngroups;
}

// ---- C phase --------------------------------------------------------------------------------------------------------------------
struct Seg4 {                     // one segment = tile (I, J), k-steps (4 points) [klo, khi) of the slab, accumulated into workspace slot `slot`;
    int I, J, n, slot;            // it touches the n chunks from chunk lo on (the first and the last one possibly in part)
    long long lo;
    int klo, khi;
    bool valid;
};
struct Slab4 {                    // what a segment needs to know about its slab
    const double* panel;
    const unsigned* gen_flags;    // generation counters of the slab's ring panel
    unsigned round;               // ... which must reach round * (generator CTAs of the block)
    long long c0;                 // first chunk of the slab
    bool first;
};

// thread 0: K_uf tiles (and weights) of chunk `cc` of the slab -> stage of pipeline position gc
template <int TM, int NB, int DPAD, bool WEIGHTED>
__device__ __forceinline__ void issue_tiles(const Params& p, const Sm4& sm, const Slab4& sl, const Seg4& sg, const long long cc, const unsigned gc) {
    using S = Smem4<TM, NB, DPAD>;
    constexpr int LDB = S::LDB;
    constexpr unsigned tile_bytes = NB * LDB * 8;
    const bool diag = sg.I == sg.J;
    const int s = gc % kKStages;
    const unsigned use = gc / kKStages;
    if (use >= 1) mbar_wait_(&sm.empty[s], (use - 1) & 1u);      // every warp has released the previous chunk in this stage
    double* st = sm.u + (S::kst - S::u) + (size_t)s * S::KSTAGE;
    mbar_expect_tx(&sm.full[s], (diag ? 1u : 2u) * tile_bytes + (WEIGHTED ? NB * 8u : 0u));
    tma_load_1d(st, sl.panel + ((size_t)cc * p.nblk + sg.I) * (size_t)(NB * LDB), tile_bytes, &sm.full[s]);
    if (!diag) tma_load_1d(st + NB * LDB, sl.panel + ((size_t)cc * p.nblk + sg.J) * (size_t)(NB * LDB), tile_bytes, &sm.full[s]);
    if (WEIGHTED) tma_load_1d(st + 2 * NB * LDB, p.w + (sl.c0 + cc) * NB, NB * 8, &sm.full[s]);
}
// thread 0: wait until the slab's rows of blocks I and J are generated, then start the segment's first loads.  Called BEFORE the
// epilogue of the previous segment so that the tiles fly while that epilogue runs.
template <int TM, int NB, int DPAD, bool WEIGHTED>
__device__ __forceinline__ void seg_prologue(const Params& p, const Sm4& sm, const Slab4& sl, const Seg4& sg, const unsigned g, long long& t_wait) {
    long long tw = 0;
    if (p.dbg) tw = clock64();
    spin_ge(sl.gen_flags + sg.I, sl.round * (unsigned)(gen_lo(sg.I + 1, p.ncta, p.nblk) - gen_lo(sg.I, p.ncta, p.nblk)));
    if (sg.I != sg.J) spin_ge(sl.gen_flags + sg.J, sl.round * (unsigned)(gen_lo(sg.J + 1, p.ncta, p.nblk) - gen_lo(sg.J, p.ncta, p.nblk)));
    if (p.dbg) t_wait += clock64() - tw;
    fence_proxy_async();
    for (int c = 0; c < kKStages - 1 && c < sg.n; ++c) issue_tiles<TM, NB, DPAD, WEIGHTED>(p, sm, sl, sg, sg.lo + c, g + c);
}


// Adds (first = false) or stores (first = true) a warp-tiled register accumulator into workspace slot `slot`.  The CTA is the slot's only
// writer and a thread always owns the same elements, so the fire-and-forget reductions (RED.ADD.F64, no round trip) arrive in slab order:
// deterministic.
template <int TM, int NT, bool DIAG>
__device__ __forceinline__ void store_acc4(const Params& p, const double (&acc)[TM / 16][TM / 32][2], const int slot, const bool first) {
    constexpr int NWARPS = NT / 32, WR = NT / 128, WM = TM / WR, WN = TM / 4, MI = WM / 8, NJ = WN / 8, SB = TM / 4, SI = SB / 8, FJ1 = 1;
    constexpr int AI = DIAG ? 2 * SI : MI, AJ = DIAG ? SI : NJ;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int wr = warp >> 2, wc = warp & 3;
    const int s0r = (0x32103321 >> (4 * warp)) & 0xf, s0c = (0x32102110 >> (4 * warp)) & 0xf;
    const int s1r = warp < SI ? 2 : 3, s1c0 = (warp % SI) * 8;
    const bool has1 = DIAG && warp < 2 * SI;
    (void)NWARPS;
    double* out = p.partial + (size_t)slot * (TM * TM);
    auto put = [&](double* q, double v0, double v1) {
        if (first) *reinterpret_cast<double2*>(q) = make_double2(v0, v1);
        else {
            asm volatile("red.relaxed.gpu.global.add.f64 [%0], %1;\n" ::"l"(q), "d"(v0) : "memory");
            asm volatile("red.relaxed.gpu.global.add.f64 [%0], %1;\n" ::"l"(q + 1), "d"(v1) : "memory");
        }
    };
    if (DIAG) {
#pragma unroll
        for (int sl = 0; sl < 2; ++sl) {
            if (sl == 1 && !has1) break;
            const int r0 = (sl ? s1r : s0r) * SB, c0 = sl ? s1c0 : s0c * SB;
#pragma unroll
            for (int i = 0; i < SI; ++i)
#pragma unroll
                for (int j = 0; j < (sl ? FJ1 : SI); ++j) {
                    const int rr = r0 + 8 * i + (lane >> 2), cc = c0 + 8 * j + 2 * (lane & 3);
                    put(out + rr * TM + cc, acc[sl * SI + i][j][0], acc[sl * SI + i][j][1]);
                }
        }
    } else {
#pragma unroll
        for (int i = 0; i < AI; ++i)
#pragma unroll
            for (int j = 0; j < AJ; ++j) {
                const int rr = wr * WM + 8 * i + (lane >> 2), cc = wc * WN + 8 * j + 2 * (lane & 3);
                put(out + rr * TM + cc, acc[i][j][0], acc[i][j][1]);
            }
    }
}

// the segment's main loop and epilogue (its first loads were started by seg_prologue); `nx` = the CTA's next segment of the slab
template <int TM, int NB, int DPAD, int NT, bool WEIGHTED, bool DIAG>
__device__ __forceinline__ void run_segment4(const Params& p, const Sm4& sm, const Slab4& sl, const Seg4& sg, const Seg4& nx, unsigned& g,
                                             long long& t_wait, double (&acc)[TM / 16][TM / 32][2], const bool zero_acc, const bool flush) {
    const bool first = sl.first;
    const int nchunks = sg.n, slot = sg.slot;
    using S = Smem4<TM, NB, DPAD>;
    constexpr int LDB = S::LDB;
    constexpr int NWARPS = NT / 32;
    constexpr int WR = NT / 128;
    constexpr int WM = TM / WR, WN = TM / 4;
    constexpr int MI = WM / 8, NJ = WN / 8;
    constexpr int KS = NB / 4;
    constexpr int SB = TM / 4, SI = SB / 8;
    static_assert(!DIAG || NWARPS == 8, "diagonal sub-block schedule is written for 8 warps");
    constexpr int FI = DIAG ? SI : MI, FJ = DIAG ? SI : NJ;
    constexpr int FJ1 = 1;                           // slot 1 of a diagonal tile: one 8-column block of a sub-block
    constexpr int AI = DIAG ? 2 * SI : MI, AJ = DIAG ? SI : NJ;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    long long t_start = 0;
    if (p.dbg) t_start = clock64();
    double* kst = sm.u + (S::kst - S::u);

    const int wr = warp >> 2, wc = warp & 3;
    const int s0r = (0x32103321 >> (4 * warp)) & 0xf, s0c = (0x32102110 >> (4 * warp)) & 0xf;
    // Diagonal tile: 10 of the 16 SB x SB sub-blocks are on or below the diagonal.  Slot 0 = one sub-block per warp; the two remaining
    // sub-blocks (2,0) and (3,0) are cut into SI column blocks each and dealt one per warp (slot 1): TM = 128 -> every warp holds
    // 16 + 4 DMMAs per k-step, so that the two warps of a scheduler stay balanced (a single warp reaches only 76 % of the DMMA rate).
    const int s1r = warp < SI ? 2 : 3, s1c0 = (warp % SI) * 8;
    const bool has1 = DIAG && warp < 2 * SI;
    const int kq = lane & 3;
    // fragment offsets inside a stage (doubles): A rows from the I tile, B rows from the J tile (diagonal: both from I)
    const int a_off = (DIAG ? s0r * SB : wr * WM) + (lane >> 2) + kq * LDB;
    const int b_off = (DIAG ? s0c * SB : NB * LDB + wc * WN) + (lane >> 2) + kq * LDB;
    const int a1_off = s1r * SB + (lane >> 2) + kq * LDB, b1_off = s1c0 + (lane >> 2) + kq * LDB;
    const int w_off = 2 * NB * LDB + kq;

    static_assert(AI == TM / 16 && AJ == TM / 32, "accumulator tile");
    if (zero_acc) {
#pragma unroll
        for (int i = 0; i < AI; ++i)
#pragma unroll
            for (int j = 0; j < AJ; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    }

#pragma unroll 1
    for (int c = 0; c < nchunks; ++c) {
        const unsigned gc = g + c;
        const int s = gc % kKStages;
        if (tid == 0 && c + kKStages - 1 < nchunks) issue_tiles<TM, NB, DPAD, WEIGHTED>(p, sm, sl, sg, sg.lo + c + kKStages - 1, gc + kKStages - 1);
        __syncwarp();
        const double* st = kst + (size_t)s * S::KSTAGE;
        mbar_wait_(&sm.full[s], (gc / kKStages) & 1u);

        double a[2][FI], b[2][FJ], a1[2][FI], b1[2][FJ1];
        auto load = [&](auto k_tag) {
            constexpr int k = decltype(k_tag)::value;
            const double* row = st + k * 4 * LDB;
#pragma unroll
            for (int ii = 0; ii < FI; ++ii) a[k & 1][ii] = row[a_off + 8 * ii];
#pragma unroll
            for (int jj = 0; jj < FJ; ++jj) b[k & 1][jj] = row[b_off + 8 * jj];
            double wn = 1.0;
            if (WEIGHTED) {
                wn = st[w_off + k * 4];
#pragma unroll
                for (int jj = 0; jj < FJ; ++jj) b[k & 1][jj] *= wn;
            }
            if (DIAG && has1) {
#pragma unroll
                for (int ii = 0; ii < FI; ++ii) a1[k & 1][ii] = row[a1_off + 8 * ii];
#pragma unroll
                for (int jj = 0; jj < FJ1; ++jj) b1[k & 1][jj] = row[b1_off + 8 * jj];
                if (WEIGHTED) {
#pragma unroll
                    for (int jj = 0; jj < FJ1; ++jj) b1[k & 1][jj] *= wn;
                }
            }
        };
        // k-steps of this chunk that belong to the segment (only its first / last chunk can be partial)
        const int kb = (c == 0) ? (sg.klo & (KS - 1)) : 0;
        const int ke = (c == nchunks - 1) ? ((sg.khi - 1) & (KS - 1)) + 1 : KS;
        auto kloop = [&](auto partial_tag) {
            constexpr bool PARTIAL = decltype(partial_tag)::value;
            load(std::integral_constant<int, 0>{});
            static_for<KS>([&](auto k_tag) {
                constexpr int k = decltype(k_tag)::value;
                if constexpr (k + 1 < KS) load(std::integral_constant<int, k + 1>{});
                if (!PARTIAL || (k >= kb && k < ke)) {
#pragma unroll
                    for (int ii = 0; ii < FI; ++ii)
#pragma unroll
                        for (int jj = 0; jj < FJ; ++jj) dmma884_nv(acc[ii][jj][0], acc[ii][jj][1], a[k & 1][ii], b[k & 1][jj]);
                    if (DIAG && has1) {
#pragma unroll
                        for (int ii = 0; ii < FI; ++ii)
#pragma unroll
                            for (int jj = 0; jj < FJ1; ++jj)
                                dmma884_nv(acc[(DIAG ? FI : 0) + ii][jj][0], acc[(DIAG ? FI : 0) + ii][jj][1], a1[k & 1][ii], b1[k & 1][jj]);
                    }
                }
            });
        };
        if (kb == 0 && ke == KS) kloop(std::false_type{});
        else kloop(std::true_type{});
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.empty[s]);      // this warp is done with the stage
    }
    g += (unsigned)nchunks;
    if (tid == 0 && nx.valid) seg_prologue<TM, NB, DPAD, WEIGHTED>(p, sm, sl, nx, g, t_wait);     // its tiles fly during the epilogue
    __syncwarp();

    if (flush) store_acc4<TM, NT, DIAG>(p, acc, slot, first);
    if (p.dbg && tid == 0) {
        const long long dt = clock64() - t_start;
        if (first) { p.dbg[4 * slot + 0] = sg.khi - sg.klo; p.dbg[4 * slot + 1] = dt; }      // k-steps (1/8 chunk)
        else { p.dbg[4 * slot + 0] += sg.khi - sg.klo; p.dbg[4 * slot + 1] += dt; }
        p.dbg[4 * slot + 2] = DIAG ? 1 : 0;
        p.dbg[4 * slot + 3] = blockIdx.x;
    }
}

// generator ownership: CTA i generates row block (i * nblk) / ncta; the CTAs of block b are [gen_lo(b), gen_lo(b + 1))
__host__ __device__ inline int gen_lo(int b, int ncta, int nblk) { return (int)(((long long)b * ncta + nblk - 1) / nblk); }

// phase 2 (after the last grid barrier): as reduce_items of sweep_kernel.cuh, with Psi1 summed over the generator CTAs of the
// row block and the scalars over all CTAs -- always in CTA order: deterministic
template <int TM, int NT>
__device__ __forceinline__ void reduce_items4(const Params& p, double* __restrict__ S_, int* __restrict__ ibuf, const int* __restrict__ hdr, long long* tr) {
    constexpr int SR = 4, STRIPES = TM / SR, NWARPS = NT / 32, LDS_ = TM + 1;
    constexpr int NBATCH = 4;                        // stripes whose loads are in flight together (S_ holds NBATCH stripes)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool packed = p.xr.nranks > 1;
    const long long tri = (long long)p.M * (p.M + 1) / 2;
    int* slots = ibuf + NWARPS;
    // hdr (shared memory, read from the plan at kernel start): {first entry, last entry + 1, and of the first entry: tile, first stripe, last stripe + 1,
    // offset of the tile's slot list, its length} -- no chain of dependent global loads after the grid barrier for the usual one-entry CTA
    const int en_lo = hdr[0], en_hi = hdr[1];
    for (int en = en_lo; en < en_hi; ++en) {
        long long r0_ = 0, r1_ = 0, r2_ = 0;
        if (p.dbg) r0_ = clock64();
        int tile, s_lo, s_hi, so, nseg;                      // this CTA's stripes of the tile; the tile's workspace slots in CTA order (host-built table)
        if (en == en_lo) { tile = hdr[2]; s_lo = hdr[3]; s_hi = hdr[4]; so = hdr[5]; nseg = hdr[6]; }
        else {
            tile = p.p2_items[3 * en]; s_lo = p.p2_items[3 * en + 1]; s_hi = p.p2_items[3 * en + 2];
            so = p.p2_tile_off[tile]; nseg = p.p2_tile_off[tile + 1] - so;
        }
        int I = 0;
        while ((I + 1) * (I + 2) / 2 <= tile) ++I;
        const int J = tile - I * (I + 1) / 2;
        __syncthreads();                                     // the previous entry is done with ibuf / S_
        for (int i = tid; i < nseg; i += NT) slots[i] = p.p2_slots[so + i];
        __syncthreads();
        if (p.dbg) { r1_ = clock64(); tr[0] += r1_ - r0_; }
        const bool diag = (I == J);
        for (int sb = s_lo; sb < s_hi; sb += NBATCH) {
            if (p.dbg) r1_ = clock64();
            const int nb = min(NBATCH, s_hi - sb);
            // Psi1 rows of the batch's stripes (diagonal tiles; warp w < SR owns row w of every stripe): a lane per generator CTA of the row block, all
            // stripes' loads in flight together and ahead of the partial-tile loads below; fixed-shape tree -> deterministic
            double p1v[NBATCH];
#pragma unroll
            for (int q = 0; q < NBATCH; ++q) p1v[q] = 0.0;
            if (diag && warp < SR) {
                const int lo = gen_lo(I, p.ncta, p.nblk), hi = gen_lo(I + 1, p.ncta, p.nblk);
#pragma unroll 2
                for (int qq = lo + lane; qq < hi; qq += 32) {
                    const double* src1 = p.psi1_partial + (size_t)qq * TM + sb * SR + warp;
#pragma unroll
                    for (int q = 0; q < NBATCH; ++q)
                        if (q < nb) p1v[q] += __ldcg(src1 + q * SR);
                }
            }
            if (tid < SR * TM / 2) {
                const int e = 2 * tid, rl = e / TM, c = e - rl * TM;
                double2 v[NBATCH];
#pragma unroll
                for (int q = 0; q < NBATCH; ++q) v[q] = make_double2(0.0, 0.0);
                const double* src = p.partial + (size_t)(sb * SR + rl) * TM + c;
#pragma unroll 8
                for (int sq = 0; sq < nseg; ++sq) {
                    const double* ps = src + (size_t)slots[sq] * (TM * TM);
#pragma unroll
                    for (int q = 0; q < NBATCH; ++q)
                        if (q < nb && (!diag || c <= (sb + q) * SR + rl)) {
                            const double2 x = __ldcg(reinterpret_cast<const double2*>(ps + (size_t)q * SR * TM));
                            v[q].x += x.x; v[q].y += x.y;
                        }
                }
#pragma unroll
                for (int q = 0; q < NBATCH; ++q) { S_[(q * SR + rl) * LDS_ + c] = v[q].x; S_[(q * SR + rl) * LDS_ + c + 1] = v[q].y; }
            }
            __syncthreads();
            if (p.dbg) { r2_ = clock64(); tr[1] += r2_ - r1_; }
            for (int q = 0; q < nb; ++q) {
                const int r0 = (sb + q) * SR;
                const double* Sq = S_ + (size_t)q * SR * LDS_;
                for (int e = tid; e < SR * TM; e += NT) {
                    {   // psi2[gi + gj*M]: SR consecutive rows = one 32-byte sector per column (multi-GPU: column gj of the packed lower triangle, to every rank)
                        const int rl = e % SR, c = e / SR, r = r0 + rl, gi = I * TM + r, gj = J * TM + c;
                        if (gi < p.M && gj < p.M && (!diag || c <= r)) {
                            if (packed) sgp_xchg::put1(p.xr, sgp_xchg::tri_col(gj, p.M) + gi - gj, Sq[rl * LDS_ + c]);
                            else {
                                p.psi2[(size_t)gi + (size_t)gj * p.M] = Sq[rl * LDS_ + c];
                                if (p.packed_out) p.packed_out[sgp_xchg::tri_col(gj, p.M) + gi - gj] = Sq[rl * LDS_ + c];
                            }
                        }
                    }
                    if (!packed) {   // mirror psi2[gj + gi*M]: consecutive threads -> consecutive columns (multi-GPU: the pull writes both halves)
                        const int rl = e / TM, c = e % TM, r = r0 + rl, gi = I * TM + r, gj = J * TM + c;
                        if (gi < p.M && gj < p.M && (diag ? c < r : true)) p.psi2[(size_t)gj + (size_t)gi * p.M] = Sq[rl * LDS_ + c];
                    }
                }
                if (diag && warp < SR) {      // Psi1 row r0 + warp (loaded at the top of the batch)
                    const int r = r0 + warp, gi = I * TM + r;
                    double v = 0.0;
#pragma unroll
                    for (int qs = 0; qs < NBATCH; ++qs) if (qs == q) v = p1v[qs];      // (static register indices)
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                    if (lane == 0 && gi < p.M) {
                        if (packed) sgp_xchg::put1(p.xr, tri + gi, v);
                        else { p.psi1[gi] = v; if (p.packed_out) p.packed_out[tri + gi] = v; }
                    }
                }
            }
            __syncthreads();                                 // S_ is reused by the next batch
            if (p.dbg) tr[2] += clock64() - r2_;
        }
        if (tile == 0 && s_lo == 0 && warp == NWARPS - 1) {      // scalars: a lane per CTA stripe, fixed-shape tree -> deterministic
            double sw = 0.0, sy = 0.0;
            for (int q = lane; q < p.ncta; q += 32) { sw += __ldcg(p.scal_partial + 2 * q); sy += __ldcg(p.scal_partial + 2 * q + 1); }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) { sw += __shfl_xor_sync(0xffffffffu, sw, o); sy += __shfl_xor_sync(0xffffffffu, sy, o); }
            if (lane == 0) {
                const double sc[4] = {p.variance * sw /* Psi0 = sum_n w_n k(x_n, x_n) */, sy /* sum_n w_n (ybar^2 + yvar) */, sw, (double)p.N};
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    if (packed) sgp_xchg::put1(p.xr, tri + p.M + q, sc[q]);
                    else { p.scal[q] = sc[q]; if (p.packed_out) p.packed_out[tri + p.M + q] = sc[q]; }
                }
            }
        }
    }
}

template <int TM, int NB, int DPAD, int NT, int KIND, bool WEIGHTED>
__global__ void __launch_bounds__(NT, 1) sweep4_kernel(const __grid_constant__ Params p) {
    using S = Smem4<TM, NB, DPAD>;
    constexpr int NWARPS = NT / 32;
    constexpr int RB = TM / (8 * NWARPS);
    extern __shared__ __align__(128) double smem[];
    Sm4 sm;
    sm.tab = smem + S::tab; sm.u = smem + S::u;
    sm.segtab = reinterpret_cast<int*>(smem + S::segtab);
    sm.full = reinterpret_cast<unsigned long long*>(smem + S::bars);
    sm.empty = sm.full + kKStages;
    sm.xfull = sm.empty + kKStages;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    long long t_k0 = 0, t_k1 = 0;
    if (p.dbg) t_k0 = clock64();
    if (tid == 0) {
        for (int s = 0; s < kKStages; ++s) { mbar_init(&sm.full[s], 1); mbar_init(&sm.empty[s], NWARPS); }
        for (int s = 0; s < kXStages; ++s) mbar_init(&sm.xfull[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    const int bcta = blockIdx.x;
    // generator ownership: CTA i generates row block (i * nblk) / ncta
    const int gblk = (int)(((long long)bcta * p.nblk) / p.ncta);
    const int glo = gen_lo(gblk, p.ncta, p.nblk), ghi = gen_lo(gblk + 1, p.ncta, p.nblk);
    const int gk = bcta - glo, gcnt = ghi - glo;
    auto first_points = [&]() {    // (thread 0) the first generator phase's raw points: their DRAM latency overlaps the set-up below
        const long long scs0 = min(p.slab_chunks, p.chunks);
        const long long c_lo = scs0 * gk / gcnt, c_hi = scs0 * (gk + 1) / gcnt;
        const int nch = (int)(c_hi - c_lo), ngr = (nch + kGC - 1) / kGC;
        for (int gi = 0; gi < kXStages - 1 && gi < ngr; ++gi) issue_points<TM, NB, DPAD, WEIGHTED>(p, sm, c_lo, nch, gi, 0u);
    };
    if (!p.ready && tid == 0) first_points();
    int* p2hdr = reinterpret_cast<int*>(smem + S::bars + 9);       // (the nine mbarriers use the first 9 of the 16 words)
    if (tid == 64) {   // this CTA's phase-2 work from the host-built plan: the dependent loads (DRAM after an L2 flush) complete under the sweep
        const int e0 = p.p2_cta_off[bcta], e1 = p.p2_cta_off[bcta + 1];
        p2hdr[0] = e0; p2hdr[1] = e1;
        if (e0 < e1) {
            const int t = p.p2_items[3 * e0];
            p2hdr[2] = t; p2hdr[3] = p.p2_items[3 * e0 + 1]; p2hdr[4] = p.p2_items[3 * e0 + 2];
            const int so = p.p2_tile_off[t];
            p2hdr[5] = so; p2hdr[6] = p.p2_tile_off[t + 1] - so;
            asm volatile("prefetch.global.L2 [%0];\n" ::"l"(p.p2_slots + so));
        }
    }
    // exp table: loads first, stores after the scalar sums below (their loads overlap)
    double tabv[SGP_EXP_TAB / NT];
#pragma unroll
    for (int i = 0; i < SGP_EXP_TAB / NT; ++i) tabv[i] = p.exptab[tid + i * NT];
    if (tid == 32) {   // this CTA's segments of a full slab (the same partition for every slab; the last slab clips the chunk ranges)
        const long long q0 = cta_pos(p.total_cost, p.ncta, bcta), q1 = cta_pos(p.total_cost, p.ncta, bcta + 1);
        long long pre = 0;
        int I = 0, J = 0, ns = 0;
        for (int t = 0; t < p.ntiles && pre < q1; ++t) {
            const int wt = (I == J) ? p.w_diag : p.w_off;
            long long lo = 0, hi = 0;
            if (pre + (long long)wt * p.slab_units + p.w_fixed > q0) seg_range(q0, q1, pre, wt, p.w_fixed, p.slab_units, lo, hi);
            if (lo < hi && ns < kMaxSeg - 4) {
                int* e = sm.segtab + 8 * ns++;
                e[0] = I; e[1] = J; e[2] = bcta + t; e[3] = (int)lo; e[4] = (int)hi;
            }
            pre += (long long)wt * p.slab_units + p.w_fixed;
            if (++J > I) { ++I; J = 0; }
        }
        sm.segtab[8 * ns] = -1;
    }
    if (p.ready) {     // the input data are still on their way from the host (sgp_sweep_psi_host launches this kernel beside its upload): wait for
                       // the word the copy engine writes after them; everything above ran under the copies
        if (tid == 0) {
            const long long t0 = clock64();
            while (sgp_xchg::ld_acquire_sys(p.ready) != p.ready_val) {
                __nanosleep(64);
                if (clock64() - t0 > (1ll << 34)) __trap();      // a lost upload fails loudly instead of hanging the GPU
            }
            fence_proxy_async();      // ... before the TMA reads of the points
            first_points();
        }
        __syncthreads();
    }
    {   // sum_n w_n and sum_n w_n (y_n^2 + yv_n) over this CTA's stripe of points
        const long long n_lo = p.N * bcta / p.ncta, n_hi = p.N * (bcta + 1) / p.ncta;
        double sw = 0.0, sy = 0.0;
        for (long long n = n_lo + tid; n < n_hi; n += NT) {
            const double wn = WEIGHTED ? p.w[n] : 1.0, yn = p.y[n], vn = p.yv ? p.yv[n] : 0.0;
            sw += wn;
            sy = fma(wn, fma(yn, yn, vn), sy);
        }
#pragma unroll
        for (int i = 0; i < SGP_EXP_TAB / NT; ++i) sm.tab[tid + i * NT] = tabv[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            sw += __shfl_xor_sync(0xffffffffu, sw, o);
            sy += __shfl_xor_sync(0xffffffffu, sy, o);
        }
        double* red = reinterpret_cast<double*>(sm.segtab) + (kMaxSeg - 4) * 4;      // the last four (unused) entries of the segment table
        if (lane == 0) { red[2 * warp] = sw; red[2 * warp + 1] = sy; }
        __syncthreads();
        if (tid == 0) {
            double a = 0.0, b = 0.0;
            for (int wq = 0; wq < NWARPS; ++wq) { a += red[2 * wq]; b += red[2 * wq + 1]; }
            p.scal_partial[2 * bcta] = a;
            p.scal_partial[2 * bcta + 1] = b;
        }
    }

    cooperative_groups::grid_group grid = cooperative_groups::this_grid();
    double psi1_acc[RB][2];
#pragma unroll
    for (int rb = 0; rb < RB; ++rb) psi1_acc[rb][0] = psi1_acc[rb][1] = 0.0;
    unsigned g = 0, xg = 0;
    long long t_gen = 0, t_wait = 0, n_gen = 0;     // instrumentation (p.dbg)
    long long t_pro = 0, t_seg = 0, t_end = 0;      // per slab: first prologue, segments, closing barrier
    // generate slab s into panel s % nring and publish it
    auto generate = [&](int s) {
        const long long slab_c0 = (long long)s * p.slab_chunks;
        const long long scs = min(p.slab_chunks, p.chunks - slab_c0);
        const int r = s % p.nring;
        double* panel = p.Kbuf + (size_t)r * (size_t)p.slab_doubles;
        unsigned* fl = p.flags + (size_t)r * (p.nblk + 1);
        long long t0 = 0;
        if (p.dbg) t0 = clock64();
        // the panel is free once every CTA has consumed the slab that lived in it (slab s - nring)
        if (s >= p.nring && tid == 0) spin_ge(fl + p.nblk, (unsigned)(s / p.nring) * (unsigned)p.ncta);
        if (p.dbg) { const long long t1 = clock64(); t_wait += t1 - t0; t0 = t1; }
        gen_phase<TM, NB, DPAD, NT, KIND, WEIGHTED>(p, sm, gblk, slab_c0 + scs * gk / gcnt, slab_c0 + scs * (gk + 1) / gcnt, slab_c0, panel, xg, psi1_acc, s == 0);
        fence_proxy_async();          // generic-proxy writes (panel in global, scratch in shared) before the async-proxy (TMA) accesses
        __threadfence();
        __syncthreads();
        if (tid == 0) red_release_inc(fl + gblk);
        if (p.dbg) { t_gen += clock64() - t0; n_gen += scs * (gk + 1) / gcnt - scs * gk / gcnt; }
    };

    if (p.dbg) t_k1 = clock64();
    double acc[TM / 16][TM / 32][2];
    const bool keep = sm.segtab[0] >= 0 && sm.segtab[8] < 0;      // exactly one segment per slab (CTA-uniform; the table is complete: barrier above)
    bool kept = false;
    generate(0);
    for (int s = 0; s < p.nslabs; ++s) {
        if (s + 1 < p.nslabs) generate(s + 1);        // one slab ahead: the consumers of slab s + 1 will not wait for it
        const long long slab_c0 = (long long)s * p.slab_chunks;
        const long long scs = min(p.slab_chunks, p.chunks - slab_c0);
        Slab4 sl;
        sl.panel = p.Kbuf + (size_t)(s % p.nring) * (size_t)p.slab_doubles;
        sl.gen_flags = p.flags + (size_t)(s % p.nring) * (p.nblk + 1);
        sl.round = (unsigned)(s / p.nring) + 1u;
        sl.c0 = slab_c0;
        sl.first = s == 0;
        // this CTA's segments of the slab, in tile order (table built at kernel start)
        int k = 0;
        auto next_seg = [&]() {
            Seg4 sg;
            sg.valid = false; sg.I = sg.J = sg.n = sg.slot = sg.klo = sg.khi = 0; sg.lo = 0;
            const int kmax = (int)(scs * (NB / 4));          // k-steps of this slab (the last slab may be shorter)
            while (k < kMaxSeg && !sg.valid) {
                const int* e = sm.segtab + 8 * k;
                if (e[0] < 0) { k = kMaxSeg; break; }
                const int lo = min(e[3], kmax), hi = min(e[4], kmax);
                if (lo < hi) {
                    sg.valid = true; sg.I = e[0]; sg.J = e[1]; sg.slot = e[2]; sg.klo = lo; sg.khi = hi;
                    sg.lo = lo / (NB / 4); sg.n = (hi - 1) / (NB / 4) - lo / (NB / 4) + 1;
                }
                ++k;
            }
            return sg;
        };
        long long tq0 = 0, tq1 = 0, tq2 = 0;
        if (p.dbg) tq0 = clock64();
        Seg4 cur = next_seg();
        if (cur.valid && tid == 0) seg_prologue<TM, NB, DPAD, WEIGHTED>(p, sm, sl, cur, g, t_wait);
        if (p.dbg) { tq1 = clock64(); t_pro += tq1 - tq0; }
        while (cur.valid) {
            const Seg4 nx = next_seg();
            // a CTA with ONE segment per slab keeps its register tile across the slabs and stores it once after the last one
            const bool zero_acc = !keep || !kept;
            if (cur.I == cur.J) run_segment4<TM, NB, DPAD, NT, WEIGHTED, true>(p, sm, sl, cur, nx, g, t_wait, acc, zero_acc, !keep);
            else run_segment4<TM, NB, DPAD, NT, WEIGHTED, false>(p, sm, sl, cur, nx, g, t_wait, acc, zero_acc, !keep);
            kept = true;
            cur = nx;
        }
        if (p.dbg) { tq2 = clock64(); t_seg += tq2 - tq1; }
        __syncthreads();              // every warp has received its last K_uf tile of the slab
        if (tid == 0) red_release_inc(p.flags + (size_t)(s % p.nring) * (p.nblk + 1) + p.nblk);
        if (p.dbg) t_end += clock64() - tq2;
    }
    if (keep && kept) {
        if (sm.segtab[0] == sm.segtab[1]) store_acc4<TM, NT, true>(p, acc, sm.segtab[2], true);
        else store_acc4<TM, NT, false>(p, acc, sm.segtab[2], true);
    }
    if (p.dbg && tid == 0) {   // records after the segment slots: {generator chunks, generator clocks, 2, cta}, {1, dependency-wait clocks, 3, cta}
        long long* d = p.dbg + 4 * (size_t)(p.ncta + p.ntiles);
        d[8 * bcta + 0] = n_gen; d[8 * bcta + 1] = t_gen; d[8 * bcta + 2] = 2; d[8 * bcta + 3] = bcta;
        d[8 * bcta + 4] = 1; d[8 * bcta + 5] = t_wait; d[8 * bcta + 6] = 3; d[8 * bcta + 7] = bcta;
    }
    long long t_k2 = 0, t_k3 = 0;
    if (p.dbg) t_k2 = clock64();
    // Psi1 rows of this CTA's generator block: sum over the eight point positions (lane / 4); lanes 0..3 hold two rows each
#pragma unroll
    for (int rb = 0; rb < RB; ++rb)
#pragma unroll
        for (int v = 0; v < 2; ++v) {
            double x = psi1_acc[rb][v];
            x += __shfl_xor_sync(0xffffffffu, x, 4);
            x += __shfl_xor_sync(0xffffffffu, x, 8);
            x += __shfl_xor_sync(0xffffffffu, x, 16);
            if (lane < 4) p.psi1_partial[(size_t)bcta * TM + warp * (8 * RB) + rb * 8 + 2 * lane + v] = x;
        }
    __threadfence();
    if (p.coop) grid.sync();
    else {   // plain launch (one CTA per SM, all resident): the last CTA to take a ticket releases everybody; a lost CTA traps after the time-out
        __syncthreads();
        if (tid == 0) {
            if (atomicAdd(p.gbar, 1u) == (unsigned)p.ncta - 1u) {
                p.gbar[0] = 0u;
                __threadfence();
                asm volatile("st.release.gpu.global.u32 [%0], %1;\n" ::"l"(p.gbar + 1), "r"(p.bar_epoch) : "memory");
            } else {
                const long long t0 = clock64();
                while (ld_acquire(p.gbar + 1) != p.bar_epoch) { if (clock64() - t0 > (1ll << 33)) __trap(); }
            }
        }
        __syncthreads();
    }
    if (p.dbg) t_k3 = clock64();
    long long tr[3] = {0, 0, 0};
    reduce_items4<TM, NT>(p, sm.u, reinterpret_cast<int*>(sm.u + 16 * (TM + 1)), p2hdr, tr);
    // the dependency counters are dead after the last grid barrier: leave them zeroed for the next launch (no memset per sweep on the host)
    if (bcta == 0)
        for (int i = tid; i < p.nring * (p.nblk + 1); i += NT) p.flags[i] = 0u;
    if (p.xr.nranks > 1) {      // sum over the ranks (xchg.cuh): two-shot all-reduce of the packed statistics over NVLink peer memory
        if (bcta == 0 && tid == 0 && ((p.M + 4 + (long long)p.M * (p.M + 1) / 2) & 1)) sgp_xchg::put1(p.xr, (long long)p.M * (p.M + 1) / 2 + p.M + 4, 0.0);
        long long tx[5];
        const long long tx0 = clock64();
        sgp_xchg::allreduce_stats(p.xr, p.stats_out, p.M, p.M + 4, bcta, p.ncta, p.dbg ? tx : nullptr);
        if (p.dbg && tid == 0) {   // {publish A, wait A, 8, reduce-scatter}, {publish B + wait B, expand, 9, phase 2 before the exchange}
            long long* d4 = p.dbg + 4 * (size_t)(p.ncta + p.ntiles) + 24 * (size_t)p.ncta;
            d4[8 * bcta + 0] = tx[0] - tx0; d4[8 * bcta + 1] = tx[1] - tx[0]; d4[8 * bcta + 2] = 8; d4[8 * bcta + 3] = tx[2] - tx[1];
            d4[8 * bcta + 4] = tx[3] - tx[2]; d4[8 * bcta + 5] = tx[4] - tx[3]; d4[8 * bcta + 6] = 9; d4[8 * bcta + 7] = tx0 - t_k3;
        }
    }
    if (p.dbg && tid == 0) {   // timeline of this CTA: {setup clocks, slab-loop clocks, 4, cta}, {final barrier clocks, phase-2 (+ exchange) clocks, 5, cta}
        long long* d = p.dbg + 4 * (size_t)(p.ncta + p.ntiles) + 8 * (size_t)p.ncta;
        d[8 * bcta + 0] = t_k1 - t_k0; d[8 * bcta + 1] = t_k2 - t_k1; d[8 * bcta + 2] = 4; d[8 * bcta + 3] = bcta;
        d[8 * bcta + 4] = t_k3 - t_k2; d[8 * bcta + 5] = clock64() - t_k3; d[8 * bcta + 6] = 5; d[8 * bcta + 7] = bcta;
        long long* d2 = d + 8 * (size_t)p.ncta;
        d2[4 * bcta + 0] = tr[0]; d2[4 * bcta + 1] = tr[1]; d2[4 * bcta + 2] = 6; d2[4 * bcta + 3] = tr[2];
        long long* d3 = d2 + 4 * (size_t)p.ncta;
        d3[4 * bcta + 0] = t_pro; d3[4 * bcta + 1] = t_seg; d3[4 * bcta + 2] = 7; d3[4 * bcta + 3] = t_end;
    }
}

template <int TM, int NB, int DPAD, int NT, int KIND>
int launch4_t(sgp_ctx* ctx, const Params& p, bool weighted, int grid) {
    using S = Smem4<TM, NB, DPAD>;
    auto kern = weighted ? sweep4_kernel<TM, NB, DPAD, NT, KIND, true> : sweep4_kernel<TM, NB, DPAD, NT, KIND, false>;
    SGP_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S::bytes));
    void* args[] = {const_cast<Params*>(&p)};
    if (p.coop) SGP_CUDA(ctx, cudaLaunchCooperativeKernel((const void*)kern, dim3(grid), dim3(NT), args, S::bytes, ctx->stream));
    else SGP_CUDA(ctx, cudaLaunchKernel((const void*)kern, dim3(grid), dim3(NT), args, S::bytes, ctx->stream));
    ctx->last_grid = grid; ctx->last_block = NT; ctx->last_smem = (int)S::bytes;
    SGP_CUDA(ctx, cudaGetLastError());
    return SGP_OK;
}

template <int TM, int NB, int NT, int KIND>
int launch4_d(sgp_ctx* ctx, const Params& p, bool weighted, int grid, int dpad) {
    switch (dpad) {
        case 4: return launch4_t<TM, NB, 4, NT, KIND>(ctx, p, weighted, grid);
        case 8: return launch4_t<TM, NB, 8, NT, KIND>(ctx, p, weighted, grid);
        default: return launch4_t<TM, NB, 16, NT, KIND>(ctx, p, weighted, grid);
    }
}

template <int KIND>
int launch4_kind(sgp_ctx* ctx, const Params& p, bool weighted, int grid, int dpad, int TM) {
    constexpr int NB = 32, NT = 256;
    if (TM == 128) return launch4_d<128, NB, NT, KIND>(ctx, p, weighted, grid, dpad);
    return launch4_d<64, NB, NT, KIND>(ctx, p, weighted, grid, dpad);
}

// defined in sweep4_se.cu / sweep4_m32.cu / sweep4_m52.cu
int launch4_se(sgp_ctx* ctx, const Params& p, bool weighted, int grid, int dpad, int TM);
int launch4_m32(sgp_ctx* ctx, const Params& p, bool weighted, int grid, int dpad, int TM);
int launch4_m52(sgp_ctx* ctx, const Params& p, bool weighted, int grid, int dpad, int TM);

}  // namespace sgp_sweep4
