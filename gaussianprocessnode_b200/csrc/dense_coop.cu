// M x M factorisations of the path as ONE cooperative kernel per job (M = 20 ... 4096: latency-bound, so the design minimises the
// critical path, not the FLOP count):
//
//   dense_job_kernel   [build A] -> blocked right-looking Cholesky, with X = L^-1 and S = X'X built by the CTAs that would otherwise wait for
//                      the diagonal-block factorisation -> [mu, Uv | transposed copy of L].  One or TWO independent jobs per launch.
//       build:  A = Lambda_prior + w Psi2 (the N-th `prod`; index-reversed, see below), A = K_uu(Z) + jitter I, A = Sigma + mu mu', or A as given
//       Cholesky, per 64-wide panel:  panel L21 = A21 Dinv'  (16 x 64 row strips, one per CTA)  | grid barrier |
//                      CTA 0 forms the next diagonal block in shared memory (A - P P' from ONE shared-memory copy of the panel block P) and
//                      factorises it there, while the other CTAs work through the step's list of 64 x 64 tile tasks: trailing update,
//                      row block k of X, the running products P(i, j) for the rows below, row block k - 1 of X into S | grid barrier
//       diagonal block (64 x 64, one CTA, shared memory): two 32 x 32 Cholesky factorisations by ONE WARP (rolled pivot loop, the row in
//                      rotating registers, finished columns read back from a column-major shared-memory copy), the forward substitution of
//                      the rows below and the inverse by two more warps one pivot behind, DMMA products for the off-diagonal parts
//       N-th prod:     the job factorises J Lambda J (J = index reversal): the inverse of THAT factor is the Cholesky factor of Sigma up to
//                      the reversal, so mu = U0' (U0 xi) and Uv = chol(Sigma + mu mu').U = G U0 with the closed-form factor G of I + p p'
//                      (a running sum down the columns of X in the kernel's tail) -- no second factorisation
//       every GEMM-shaped piece is a `gemm_task`: global -> registers -> shared-memory staging (software pipelined), DMMA.8x8x4
//
// Replaces LAPACK potrf / potri behind fastcholesky! / cholinv and the posterior update of the N-th `prod`
// (GPnode/UniSGPnode.jl:62-73; regression_kin40k.ipynb:183-184; Pendulum_Wishart_2d.ipynb:2542-2543).
#include "sgp_internal.cuh"
#include <cooperative_groups.h>
#include <algorithm>
#include <cstdlib>

namespace cg = cooperative_groups;

int sgp_gemm(sgp_ctx* ctx, int opA, int opB, int m, int n, int k, double alpha, const double* A, int lda, const double* B, int ldb,
             double beta, double* C, int ldc, int lower_only);

namespace {

constexpr int TB = 64;        // panel width = diagonal block
constexpr int CT = 256;       // threads per CTA (8 warps)
constexpr int STAGE_A = TB * (64 + 4), STAGE_B = 64 * (TB + 4);          // staging of the largest tile configurations (Tile::LDA/LDB)
constexpr int LDT = TB + 4;   // leading dimension of the 64 x 64 shared-memory blocks (= 4 mod 16: conflict-free DMMA fragment loads)
static_assert(4 * SGP_FLIP_UV_MAX_M <= 2 * 64 * 68 + 66 + 2 * 64 * 68, "p, G and the scan buffers in shared memory");
constexpr int kFineTiles = 148;   // S tiles of the last step / final update as 32 x 32 when 4 x their number fits about one round of a full grid (a function of M only: results do not depend on the grid)
constexpr int kPParts = 4;    // partial sums per row strip of p = X xi (more CTAs on a memory-latency-bound pass)
constexpr int SMEM_DOUBLES = 2 * TB * LDT + TB + 2 + STAGE_A + STAGE_B;     // T | Xi | rdiag | progress word | As | Bs

// ---- tile GEMM task: acc (+)= sum_{k < K} A(r, k) B(k, c) for a TR x TC tile, 8 warps arranged WR x (8 / WR) ------------------------
template <int TR, int TC, int WR, int KCH>     // KCH = K chunk staged in shared memory per barrier pair
struct Tile {
    static constexpr int WC = 8 / WR, WTR = TR / WR, WTC = TC / WC, MI = WTR / 8, NJ = WTC / 8, KC = KCH;
    static constexpr int LDA = KC + 4, LDB = TC + 4;              // both = 4 (mod 16)
    static constexpr int AP = TR * KC / CT, BP = KC * TC / CT;   // staged elements per thread
    static_assert(WTR % 8 == 0 && WTC % 8 == 0 && AP >= 1 && BP >= 1, "tile shape");
    static_assert(TR * LDA <= STAGE_A && KC * LDB <= STAGE_B, "staging area");
    struct Acc { double v[MI][NJ][2]; };
};

template <class TL>
__device__ __forceinline__ void acc_zero(typename TL::Acc& a) {
#pragma unroll
    for (int i = 0; i < TL::MI; ++i)
#pragma unroll
        for (int j = 0; j < TL::NJ; ++j) a.v[i][j][0] = a.v[i][j][1] = 0.0;
}

__device__ __forceinline__ double lds_f64(unsigned addr) {     // shared-memory load that keeps its place among the other ordered asm statements
    double v;
    asm volatile("ld.shared.f64 %0, [%1];\n" : "=d"(v) : "r"(addr) : "memory");
    return v;
}

// Staging map of a TR x KC (KC x TC) operand chunk: element e = tid + q * 256 of the chunk -> (r, kk), 8 consecutive threads along the operand's
// contiguous index (64-byte global runs), the next 4 along the other (<= 2-way bank conflicts in shared memory for either orientation).
// With 256 threads and 2, 4 or 8 groups of 8 along the fast index the map is AFFINE in q -- (r, kk) = (r0 + q dr, kk0 + q dk) -- so a thread
// keeps one base pointer per operand and a constant stride (no per-element 64-bit index arithmetic in the chunk loop), and full tiles take
// a path without bounds predicates.
template <class TL>
struct StageMap {
    static constexpr int TR = TL::WTR * (8 / TL::WC), TC = TL::WTC * TL::WC, KC = TL::KC;
    static_assert(8 % (TR / 8) == 0 && 8 % (TC / 8) == 0 && 8 % (KC / 8) == 0, "affine staging map");
    const double* pa; const double* pb;       // element q = 0 of chunk 0
    long long sa, sb, ca, cb;                 // pointer strides per q and per chunk
    int ar0, akk0, adr, adk, br0, bkk0, bdr, bdk;      // (r, kk) of q = 0 and the step per q (b: r means the column c)
    int rows, cols, K;
    bool full;
    __device__ __forceinline__ static void dec(int e, bool fast_first, int n_first, int& first, int& kk) {
        // fast_first: the tile index (r or c, n_first of them) is the contiguous one; else kk (KC of them) is
        const int lo = e & 7, mid = (e >> 3) & 3, rest = e >> 5;
        if (fast_first) { first = lo + 8 * (rest % (n_first / 8)); kk = mid + 4 * (rest / (n_first / 8)); }
        else { kk = lo + 8 * (rest % (KC / 8)); first = mid + 4 * (rest / (KC / 8)); }
    }
    __device__ __forceinline__ StageMap(const double* A, size_t a_rs, size_t a_ks, int rows_, const double* B, size_t b_ks, size_t b_cs, int cols_, int K_)
        : rows(rows_), cols(cols_), K(K_) {
        const int tid = threadIdx.x;
        int r1, k1;
        dec(tid, a_rs == 1, TR, ar0, akk0); dec(tid + CT, a_rs == 1, TR, r1, k1); adr = r1 - ar0; adk = k1 - akk0;
        dec(tid, b_cs == 1, TC, br0, bkk0); dec(tid + CT, b_cs == 1, TC, r1, k1); bdr = r1 - br0; bdk = k1 - bkk0;
        pa = A + (long long)ar0 * (long long)a_rs + (long long)akk0 * (long long)a_ks;
        pb = B + (long long)bkk0 * (long long)b_ks + (long long)br0 * (long long)b_cs;
        sa = (long long)adr * (long long)a_rs + (long long)adk * (long long)a_ks;
        sb = (long long)bdk * (long long)b_ks + (long long)bdr * (long long)b_cs;
        ca = (long long)KC * (long long)a_ks; cb = (long long)KC * (long long)b_ks;
        full = rows == TR && cols == TC && K % KC == 0;
    }
    __device__ __forceinline__ void gload(double (&ra)[TL::AP], double (&rb)[TL::BP], int k0) const {
        const double* a = pa + (long long)(k0 / KC) * ca; const double* b = pb + (long long)(k0 / KC) * cb;
        if (full) {
#pragma unroll
            for (int q = 0; q < TL::AP; ++q) ra[q] = a[q * sa];
#pragma unroll
            for (int q = 0; q < TL::BP; ++q) rb[q] = b[q * sb];
        } else {
#pragma unroll
            for (int q = 0; q < TL::AP; ++q) ra[q] = (ar0 + q * adr < rows && k0 + akk0 + q * adk < K) ? a[q * sa] : 0.0;
#pragma unroll
            for (int q = 0; q < TL::BP; ++q) rb[q] = (br0 + q * bdr < cols && k0 + bkk0 + q * bdk < K) ? b[q * sb] : 0.0;
        }
    }
    __device__ __forceinline__ void sstore(const double (&ra)[TL::AP], const double (&rb)[TL::BP], double* __restrict__ As, double* __restrict__ Bs) const {
        double* a = As + ar0 * TL::LDA + akk0; double* b = Bs + bkk0 * TL::LDB + br0;
        const int ia = adr * TL::LDA + adk, ib = bdk * TL::LDB + bdr;
#pragma unroll
        for (int q = 0; q < TL::AP; ++q) a[q * ia] = ra[q];
#pragma unroll
        for (int q = 0; q < TL::BP; ++q) b[q * ib] = rb[q];
    }
};

// A(r, k) = A[r * a_rs + k * a_ks] (r < rows), B(k, c) = B[k * b_ks + c * b_cs] (c < cols); out-of-range elements read as zero.
// Staging map: a thread owns an (8 fast x 4 slow)-interleaved element so that both the global loads (64-byte runs along the fast
// index) and the shared-memory stores (two-way bank conflicts at most, for either orientation) are efficient.
// Global loads of chunk k+1 are in flight while chunk k is multiplied.  All 256 threads must call it.
template <class TL>
__device__ __forceinline__ void gemm_task(typename TL::Acc& acc, const double* __restrict__ A, size_t a_rs, size_t a_ks, int rows,
                                          const double* __restrict__ B, size_t b_ks, size_t b_cs, int cols, int K, double* __restrict__ As,
                                          double* __restrict__ Bs, long long* __restrict__ tk = nullptr) {
    constexpr int TR = TL::WTR * (8 / TL::WC), TC = TL::WTC * TL::WC, KC = TL::KC;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wr = warp / TL::WC, wc = warp % TL::WC;
    double ra[TL::AP], rb[TL::BP];
    StageMap<TL> sm_(A, a_rs, a_ks, rows, B, b_ks, b_cs, cols, K);
    auto gload = [&](int k0) { sm_.gload(ra, rb, k0); };
    auto sstore = [&]() { sm_.sstore(ra, rb, As, Bs); };
    typename TL::Acc part[4];                              // (small warp tiles only)
    if constexpr (TL::MI * TL::NJ <= 2) {
#pragma unroll
        for (int q = 0; q < 4; ++q) acc_zero<TL>(part[q]);
    }
    long long q0 = 0, q1;
    if (tk) q0 = clock64();
#define TCLK(i) do { if (tk) { q1 = clock64(); tk[i] += q1 - q0; q0 = q1; } } while (0)      /* tk: thread-local counters */
    gload(0);
    for (int k0 = 0; k0 < K; k0 += KC) {
        __syncthreads();                       // the previous chunk's fragments have been read
        TCLK(0);                               // (tuning aid) barrier wait
        sstore();
        __syncthreads();
        TCLK(1);                               // global-load wait + staging stores
        if (k0 + KC < K) gload(k0 + KC);
        if constexpr (TL::MI * TL::NJ <= 2) {
            // Small warp tiles are shared-memory bound (384 B of fragments per DMMA): groups of four k-steps, the loads of group g + 1 issued
            // (ordered asm) before the DMMAs of group g, so that the warps' load and DMMA phases overlap instead of alternating in lock step.
            constexpr int NG = KC / 16;
            double fa[2][4][TL::MI], fb[2][4][TL::NJ];
            const unsigned a0 = (unsigned)__cvta_generic_to_shared(As + (wr * TL::WTR + (lane >> 2)) * TL::LDA + (lane & 3));
            const unsigned b0 = (unsigned)__cvta_generic_to_shared(Bs + (lane & 3) * TL::LDB + wc * TL::WTC + (lane >> 2));
            auto lgroup = [&](int g, int buf) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int ks = 4 * g + q;
#pragma unroll
                    for (int i = 0; i < TL::MI; ++i) fa[buf][q][i] = lds_f64(a0 + 8u * (unsigned)(8 * i * TL::LDA + ks * 4));
#pragma unroll
                    for (int j = 0; j < TL::NJ; ++j) fb[buf][q][j] = lds_f64(b0 + 8u * (unsigned)(ks * 4 * TL::LDB + 8 * j));
                }
            };
            lgroup(0, 0);
#pragma unroll
            for (int g = 0; g < NG; ++g) {
                if (g + 1 < NG) lgroup(g + 1, (g + 1) & 1);
                // (four accumulator sets, k-steps mod 4: a dependent DMMA waits long when eight warps share the pipe)
#pragma unroll
                for (int q = 0; q < 4; ++q)
#pragma unroll
                    for (int i = 0; i < TL::MI; ++i)
#pragma unroll
                        for (int j = 0; j < TL::NJ; ++j) dmma884(part[q].v[i][j][0], part[q].v[i][j][1], fa[g & 1][q][i], fb[g & 1][q][j]);
            }
        } else {
#pragma unroll
            for (int ks = 0; ks < KC / 4; ++ks) {
                double a[TL::MI], b[TL::NJ];
#pragma unroll
                for (int i = 0; i < TL::MI; ++i) a[i] = As[(wr * TL::WTR + 8 * i + (lane >> 2)) * TL::LDA + ks * 4 + (lane & 3)];
#pragma unroll
                for (int j = 0; j < TL::NJ; ++j) b[j] = Bs[(ks * 4 + (lane & 3)) * TL::LDB + wc * TL::WTC + 8 * j + (lane >> 2)];
#pragma unroll
                for (int i = 0; i < TL::MI; ++i)
#pragma unroll
                    for (int j = 0; j < TL::NJ; ++j) dmma884(acc.v[i][j][0], acc.v[i][j][1], a[i], b[j]);
            }
        }
        TCLK(2);                               // fragment loads + DMMA
    }
    if constexpr (TL::MI * TL::NJ <= 2) {
#pragma unroll
        for (int i = 0; i < TL::MI; ++i)
#pragma unroll
            for (int j = 0; j < TL::NJ; ++j)
#pragma unroll
                for (int h = 0; h < 2; ++h) acc.v[i][j][h] += (part[0].v[i][j][h] + part[1].v[i][j][h]) + (part[2].v[i][j][h] + part[3].v[i][j][h]);
    }
#undef TCLK
}

// The same for the plain GEMMs (long K, one tile per CTA): the staging area is DOUBLE-BUFFERED, so a warp that has finished the DMMAs of chunk k
// stores chunk k + 1 into the other buffer while the slower warps still multiply -- one barrier per chunk instead of two, and no
// store phase in which the DMMA pipe idles.  buf: 2 x (TR x LDA + KC x LDB) doubles.  Warp tiles of 16 x 16 and larger only.
template <class TL>
__device__ __forceinline__ void gemm_task_db(typename TL::Acc& acc, const double* __restrict__ A, size_t a_rs, size_t a_ks, int rows,
                                             const double* __restrict__ B, size_t b_ks, size_t b_cs, int cols, int K, double* __restrict__ buf) {
    static_assert(TL::MI * TL::NJ > 2, "large warp tiles only");
    constexpr int TR = TL::WTR * (8 / TL::WC), TC = TL::WTC * TL::WC, KC = TL::KC, SA = TR * TL::LDA, SB = KC * TL::LDB;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wr = warp / TL::WC, wc = warp % TL::WC;
    double ra[TL::AP], rb[TL::BP];
    StageMap<TL> sm_(A, a_rs, a_ks, rows, B, b_ks, b_cs, cols, K);
    auto gload = [&](int k0) { sm_.gload(ra, rb, k0); };
    auto sstore = [&](double* __restrict__ As, double* __restrict__ Bs) { sm_.sstore(ra, rb, As, Bs); };
    gload(0);
    sstore(buf, buf + SA);
    __syncthreads();
    int cur = 0;
    for (int k0 = 0; k0 < K; k0 += KC) {
        const double* As = buf + cur * (SA + SB); const double* Bs = As + SA;
        const bool more = k0 + KC < K;
        if (more) gload(k0 + KC);
#pragma unroll
        for (int ks = 0; ks < KC / 4; ++ks) {
            double a[TL::MI], b[TL::NJ];
#pragma unroll
            for (int i = 0; i < TL::MI; ++i) a[i] = As[(wr * TL::WTR + 8 * i + (lane >> 2)) * TL::LDA + ks * 4 + (lane & 3)];
#pragma unroll
            for (int j = 0; j < TL::NJ; ++j) b[j] = Bs[(ks * 4 + (lane & 3)) * TL::LDB + wc * TL::WTC + 8 * j + (lane >> 2)];
#pragma unroll
            for (int i = 0; i < TL::MI; ++i)
#pragma unroll
                for (int j = 0; j < TL::NJ; ++j) dmma884(acc.v[i][j][0], acc.v[i][j][1], a[i], b[j]);
        }
        if (more) {
            double* An = buf + (cur ^ 1) * (SA + SB);
            sstore(An, An + SA);           // (the other buffer was last read before the previous barrier)
        }
        __syncthreads();
        cur ^= 1;
    }
}

// C(r, c) = alpha acc + beta C(r, c), r < rows, c < cols; C column-major (ldc).  lower_only: only elements with grow0 + r >= gcol0 + c.
// Ct (optional): the same value to Ct(c, r) (mirror of a symmetric result).
template <class TL>
__device__ __forceinline__ void store_task(const typename TL::Acc& acc, double* __restrict__ C, int ldc, int rows, int cols, double alpha,
                                           double beta, double* __restrict__ Ct = nullptr, bool lower_only = false, int grow0 = 0,
                                           int gcol0 = 0) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int wr = warp / TL::WC, wc = warp % TL::WC;
#pragma unroll
    for (int i = 0; i < TL::MI; ++i)
#pragma unroll
        for (int j = 0; j < TL::NJ; ++j)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int r = wr * TL::WTR + 8 * i + (lane >> 2), c = wc * TL::WTC + 8 * j + 2 * (lane & 3) + h;
                if (r < rows && c < cols && (!lower_only || grow0 + r >= gcol0 + c)) {
                    double* p = C + (size_t)r + (size_t)c * ldc;
                    const double v = fma(alpha, acc.v[i][j][h], beta == 0.0 ? 0.0 : beta * *p);
                    *p = v;
                    if (Ct) Ct[(size_t)c + (size_t)r * ldc] = v;
                }
            }
}

using T64 = Tile<64, 64, 4, 32>;      // warp tile 16 x 32
using T32 = Tile<32, 32, 4, 64>;      // warp tile  8 x 16: little MMA work per chunk, so a long chunk per barrier pair
using T16 = Tile<16, 64, 2, 64>;      // row strip: warp tile 8 x 16
using T63 = Tile<64, 32, 4, 64>;      // 64 x 32: warp tile 16 x 16, long chunks (the plain GEMMs: K is long)
using TX = Tile<64, 16, 8, 64>;       // column slice of a row block of X: warp tile 8 x 16

// The 32 x 32 sub-tile task every phase is made of, as ONE non-inlined routine (the kernel's instruction footprint decides its speed: each
// phase runs its code only a few times per launch).  C = alpha A B + beta C with the masks of store_task.
struct Task32 {
    const double* A; size_t a_rs, a_ks; int rows;
    const double* B; size_t b_ks, b_cs; int cols; int K;
    double* C; int ldc; double alpha, beta; double* Ct; bool lower_only; int grow0, gcol0;
    long long* tk;
};
// ... and the 64 x 64 tile form (a quarter of the operand traffic per flop: what the worker CTAs run beside CTA 0's factorisations)
__device__ __noinline__ void run_task64(const Task32& t, double* __restrict__ As, double* __restrict__ Bs) {
    T64::Acc acc; acc_zero<T64>(acc);
    gemm_task<T64>(acc, t.A, t.a_rs, t.a_ks, t.rows, t.B, t.b_ks, t.b_cs, t.cols, t.K, As, Bs, t.tk);
    if (t.tk) { t.tk[3] += 1; t.tk[4] += (t.K + T64::KC - 1) / T64::KC; }
    store_task<T64>(acc, t.C, t.ldc, t.rows, t.cols, t.alpha, t.beta, t.Ct, t.lower_only, t.grow0, t.gcol0);
}
__device__ __noinline__ void run_task32(const Task32& t, double* __restrict__ As, double* __restrict__ Bs) {
    T32::Acc acc; acc_zero<T32>(acc);
    gemm_task<T32>(acc, t.A, t.a_rs, t.a_ks, t.rows, t.B, t.b_ks, t.b_cs, t.cols, t.K, As, Bs, t.tk);
    if (t.tk) { t.tk[3] += 1; t.tk[4] += (t.K + T32::KC - 1) / T32::KC; }
    store_task<T32>(acc, t.C, t.ldc, t.rows, t.cols, t.alpha, t.beta, t.Ct, t.lower_only, t.grow0, t.gcol0);
}

// ---- the 64 x 64 diagonal block: factor + inverse in shared memory (leading dimension LDT), one CTA ------------------------------------
// Measured on B200 (tools/lat_microbench.cu, tools/dense_clocks.py): a DFMA chain costs 8.7 clocks per link, rsqrt 66, a double shuffle 30
// with an issue cost that makes 62 shuffles per pivot ~700 clocks, a shared-memory load ~4 issue clocks (29 latency), a dependent DMMA 26.
// And: fully unrolled pivot loops (~50 KB of straight-line code per copy) run from instruction-cache misses at ~4 000 clocks per pivot.
// Hence: ROLLED loops over the pivots, the not-yet-final part of a row / column in registers that ROTATE by one per pivot (all register
// indices static), finished columns published through a column-major copy `Lc` in shared memory and read back as aligned 16-byte broadcasts
// (no shuffles), and the next pivot's reciprocal square root issued before the current pivot's column updates.
// NU = columns updated per pivot: 31 for the first 16 pivots, 15 for the last 16 (736 instead of 496 updates; loop bodies of ~80 / ~45
// instructions).

// Progress of the pivot loop, published by the factorising warp and polled by the warps that consume the finished columns one pivot behind
// (forward substitution of the rows below, inverse): release / acquire at CTA scope on a shared-memory word.
__device__ __forceinline__ void prog_publish(int* prog, int v) {
    asm volatile("st.release.cta.shared.u32 [%0], %1;\n" ::"r"((unsigned)__cvta_generic_to_shared(prog)), "r"(v) : "memory");
}
__device__ __forceinline__ void prog_wait(const int* prog, int above) {      // until *prog > above
    const unsigned a = (unsigned)__cvta_generic_to_shared(prog);
    int v;
    do { asm volatile("ld.acquire.cta.shared.u32 %0, [%1];\n" : "=r"(v) : "r"(a) : "memory"); } while (v <= above);
}

constexpr int LCS = 64;                                          // stride of Lc: Lc[j * LCS + ((j + 1) & 1) + i] = L[i][j]
__device__ __forceinline__ int lc_col(int j) { return j * LCS + ((j + 1) & 1); }      // ... so that &Lc[.. + j + 1] is 16-byte aligned

// a[t - 1] = a[t] - l * colp[t - 1], t = 1 .. NU   (colp = &L[j + 1][j] in Lc; entries past row 31 are garbage that lands in dead slots)
template <int NU>
__device__ __forceinline__ void rot_update(double (&a)[32], double l, const double* __restrict__ colp) {
#pragma unroll
    for (int t = 1; t + 1 <= NU; t += 2) {
        const double2 v = *reinterpret_cast<const double2*>(colp + (t - 1));
        a[t - 1] = fma(-l, v.x, a[t]);
        a[t] = fma(-l, v.y, a[t + 1]);
    }
    if (NU & 1) a[NU - 1] = fma(-l, colp[NU - 1], a[NU]);
}

// 1 / sqrt(d) for a positive normal d, branch-free: MUFU.RSQ64H seed (2^-22) and one third-order correction -- 6 instructions that ptxas can
// interleave with the column updates (the library rsqrt carries a slow-path branch that splits the basic block).
__device__ __forceinline__ double rsqrt_pos(double d) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;\n" : "=d"(y) : "d"(d));
    const double t = d * y;
    const double e = fma(-t, y, 1.0);                      // 1 - d y^2
    const double p = fma(0.375, e, 0.5);
    return fma(y * e, p, y);                               // y (1 + e / 2 + 3 e^2 / 8)
}

template <int NU>
__device__ __forceinline__ void chol_pivot(double (&a)[32], double& dg, double& d, double& r, int& bad, int j, int lane, double* __restrict__ row,
                                           double* __restrict__ Lc, double* __restrict__ rdiag_off, int* __restrict__ prog) {
    double l = a[0] * r;
    if (lane == j) { l = d * r; rdiag_off[j] = r; }
    dg = fma(-l, l, dg);                                   // this lane's own diagonal element (meaningful for lanes > j)
    double dn = __shfl_sync(0xffffffffu, dg, (j + 1) & 31);       // the next pivot
    row[j] = (lane >= j) ? l : 0.0;                       // column j of L is final (in place in T)
    const int cj = lc_col(j);
    Lc[cj + lane] = l;
    __syncwarp();
    if (lane == 0) prog_publish(prog, j + 1);             // column j (and 1 / L_jj) are out: the lagging warps may use them
    // one basic block from here: the next pivot's scale (a chain of ~60 clocks) is interleaved with the column updates
    const bool neg = !(dn > 0.0) && j < 31;
    bad = (neg && bad < 0) ? j + 1 : bad;                  // first non-positive pivot (recorded once, after the loop)
    dn = neg ? 1.0 : dn;
    const double rn = rsqrt_pos(dn);
    rot_update<NU>(a, l, Lc + cj + j + 1);
    d = dn; r = rn;
}
// In-place lower Cholesky of the 32 x 32 block of T at `off` by ONE warp (lane = row); Lc receives the factor column by column,
// rdiag[off + j] = 1 / L_jj.  A non-positive pivot is recorded (rows below `nvalid` only) and replaced by 1.
__device__ __noinline__ void chol32_warp(double* __restrict__ T, int off, double* __restrict__ Lc, double* __restrict__ rdiag, int* __restrict__ info,
                                         int row0, int nvalid, int* __restrict__ prog) {
    const int lane = threadIdx.x & 31;
    double a[32];
    double* row = T + (off + lane) * LDT + off;
#pragma unroll
    for (int c = 0; c < 32; ++c) a[c] = row[c];
    double dg = row[lane];
    double d = __shfl_sync(0xffffffffu, dg, 0);
    int bad = -1;
    if (!(d > 0.0)) { bad = 0; d = 1.0; }
    double r = rsqrt_pos(d);
#pragma unroll 1
    for (int j = 0; j < 16; ++j) chol_pivot<31>(a, dg, d, r, bad, j, lane, row, Lc, rdiag + off, prog);
#pragma unroll 1
    for (int j = 16; j < 32; ++j) chol_pivot<15>(a, dg, d, r, bad, j, lane, row, Lc, rdiag + off, prog);
    if (lane == 0 && bad >= 0 && off + bad < nvalid) atomicCAS(info, 0, row0 + off + bad + 1);
}

// Rows 32..63 of the panel: L21 = A21 L11^-T by ONE warp (lane = row; forward substitution with the rotating registers), in place in T.
__device__ __noinline__ void trsm32_rows_warp(double* __restrict__ T, const double* __restrict__ Lc, const double* __restrict__ rdiag,
                                              const int* __restrict__ prog) {
    const int lane = threadIdx.x & 31;
    double a[32];
    double* row = T + (32 + lane) * LDT;
#pragma unroll
    for (int c = 0; c < 32; ++c) a[c] = row[c];
#pragma unroll 1
    for (int j = 0; j < 16; ++j) { prog_wait(prog, j); const double l = a[0] * rdiag[j]; row[j] = l; rot_update<31>(a, l, Lc + lc_col(j) + j + 1); }
#pragma unroll 1
    for (int j = 16; j < 32; ++j) { prog_wait(prog, j); const double l = a[0] * rdiag[j]; row[j] = l; rot_update<15>(a, l, Lc + lc_col(j) + j + 1); }
}

// Xi block at `off` = inverse of the lower-triangular 32 x 32 block whose columns are in Lc, by ONE warp: lane c solves L x = e_c.
__device__ __noinline__ void inv32_warp(const double* __restrict__ Lc, int off, const double* __restrict__ rdiag, double* __restrict__ Xi,
                                        const int* __restrict__ prog) {
    const int lane = threadIdx.x & 31;
    double x[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) x[i] = (i == lane) ? 1.0 : 0.0;
    double* Xb = Xi + off * LDT + off;
#pragma unroll 1
    for (int l = 0; l < 16; ++l) {
        prog_wait(prog, l);
        const double xl = x[0] * rdiag[off + l];
        Xb[l * LDT + lane] = (l >= lane) ? xl : 0.0;      // row l of the inverse is final
        rot_update<31>(x, xl, Lc + lc_col(l) + l + 1);
    }
#pragma unroll 1
    for (int l = 16; l < 32; ++l) {
        prog_wait(prog, l);
        const double xl = x[0] * rdiag[off + l];
        Xb[l * LDT + lane] = (l >= lane) ? xl : 0.0;
        rot_update<15>(x, xl, Lc + lc_col(l) + l + 1);
    }
}

// 32 x 32 x 32 product on shared-memory operands by all 8 warps (warp tile 8 x 16): returns the warp's fragment of
//   sum_k A(r, k) B(k, c),  A(r, k) = A[r * a_rs + k * a_ks],  B(k, c) = B[k * b_ks + c * b_cs]
struct Frag32 { double v[2][2]; };
__device__ __forceinline__ Frag32 smem_mma32(const double* __restrict__ A, int a_rs, int a_ks, const double* __restrict__ B, int b_ks, int b_cs) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int wr = warp >> 1, wc = warp & 1;
    Frag32 f, g[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) g[q].v[0][0] = g[q].v[0][1] = g[q].v[1][0] = g[q].v[1][1] = 0.0;
    double a[8], b[8][2];
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) {
        const int k = ks * 4 + (lane & 3);
        a[ks] = A[(wr * 8 + (lane >> 2)) * a_rs + k * a_ks];
#pragma unroll
        for (int j = 0; j < 2; ++j) b[ks][j] = B[k * b_ks + (wc * 16 + 8 * j + (lane >> 2)) * b_cs];
    }
#pragma unroll
    for (int ks = 0; ks < 8; ++ks)           // four accumulator sets: a dependent DMMA waits ~140 clocks when eight warps share the pipe
#pragma unroll
        for (int j = 0; j < 2; ++j) dmma884(g[ks & 3].v[j][0], g[ks & 3].v[j][1], a[ks], b[ks][j]);
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int h = 0; h < 2; ++h) f.v[j][h] = (g[0].v[j][h] + g[1].v[j][h]) + (g[2].v[j][h] + g[3].v[j][h]);
    return f;
}
// C(r, c) = alpha f + beta C(r, c) for the warp's fragment; C[r * LDT + c]
__device__ __forceinline__ void smem_store32(const Frag32& f, double* __restrict__ C, double alpha, double beta) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int wr = warp >> 1, wc = warp & 1;
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            double* p = C + (wr * 8 + (lane >> 2)) * LDT + wc * 16 + 8 * j + 2 * (lane & 3) + h;
            *p = fma(alpha, f.v[j][h], beta == 0.0 ? 0.0 : beta * *p);
        }
}

// Element e of a 64 x 64 block that is column-major in global memory and row-major (LDT) in shared memory -> (r, c): 8 consecutive threads
// along a column (64-byte runs in global memory), the next 4 along the row: two-way bank conflicts at most on the shared-memory side
// (a plain r = e % 64 map is eight-way conflicted there).
__device__ __forceinline__ void blk_rc(int e, int& r, int& c) {
    const int lo = e & 7, mid = (e >> 3) & 3, rest = e >> 5;
    r = lo + 8 * (rest & 7); c = mid + 4 * (rest >> 3);
}
// Factor the nb x nb diagonal block held in T (lower part valid, identity-padded beyond nb, zeros above the diagonal), write L to A
// (global, lower part) and its inverse to Dinv (64 x 64 column-major; only the lower triangle is written).  One CTA, all 256 threads.
//   warp 0: Cholesky of the leading 32 x 32 block; one pivot behind it warp 1: L21 = A21 L11^-T (forward substitution), warp 2: X11 = L11^-1
//   all: A22 -= L21 L21', W = L21 X11 (DMMA)  |  warp 0: Cholesky of A22, warp 1 one pivot behind: X22  |  all: X21 = -X22 W
__device__ __noinline__ void factor_diag_smem(double* __restrict__ T, double* __restrict__ Xi, double* __restrict__ rdiag, double* __restrict__ Lc,
                                              double* __restrict__ A, int lda, int nb, int row0, double* __restrict__ Dinv, int* __restrict__ info,
                                              long long* __restrict__ st = nullptr) {
    const int tid = threadIdx.x, warp = tid >> 5;
    double* T21 = T + 32 * LDT; double* T22 = T + 32 * LDT + 32;
    double* X21 = Xi + 32 * LDT;
    long long s0 = 0, s1;
#define FCLK(i) do { if (st) { s1 = clock64(); st[i] += s1 - s0; s0 = s1; } } while (0)
    int* prog = reinterpret_cast<int*>(rdiag + TB);        // progress word of the pivot loop (behind the 64 reciprocal diagonals)
    if (tid == 0) *prog = 0;
    __syncthreads();
    if (st) s0 = clock64();
    // leading block: warp 0 factorises; one pivot behind, warp 1 substitutes the rows below (L21) and warp 2 builds the inverse X11
    if (warp == 0) chol32_warp(T, 0, Lc, rdiag, info, row0, nb, prog);
    else if (warp == 1) trsm32_rows_warp(T, Lc, rdiag, prog);
    else if (warp == 2) inv32_warp(Lc, 0, rdiag, Xi, prog);
    __syncthreads();
    FCLK(0);
    {   // A22 -= L21 L21' ;  W = L21 X11 (kept in the X21 slot)
        const Frag32 f = smem_mma32(T21, LDT, 1, T21, 1, LDT);
        const Frag32 g = smem_mma32(T21, LDT, 1, Xi, LDT, 1);
        smem_store32(f, T22, -1.0, 1.0);
        smem_store32(g, X21, 1.0, 0.0);
        if (tid == 0) *prog = 0;
    }
    __syncthreads();
    FCLK(2);
    // trailing block: warp 0 factorises, warp 1 builds the inverse X22 one pivot behind
    if (warp == 0) chol32_warp(T, 32, Lc, rdiag, info, row0, nb, prog);
    else if (warp == 1) inv32_warp(Lc, 32, rdiag, Xi, prog);
    __syncthreads();
    FCLK(0);
    {   // X21 = -X22 W  (in place: all reads before the writes)
        const Frag32 f = smem_mma32(Xi + 32 * LDT + 32, LDT, 1, X21, LDT, 1);
        __syncthreads();
        smem_store32(f, X21, -1.0, 0.0);
    }
    __syncthreads();
    FCLK(2);
    for (int e = tid; e < TB * TB; e += CT) {           // (the strict upper triangle of every Dinv block is zero from its allocation;
        int r, c; blk_rc(e, r, c);                      //  the grid barrier that follows publishes the stores)
        if (c <= r) {
            if (r < nb) A[(size_t)r + (size_t)c * lda] = T[r * LDT + c];
            Dinv[(size_t)r + (size_t)c * TB] = Xi[r * LDT + c];
        }
    }
    __syncthreads();
    FCLK(3);
#undef FCLK
}

// T <- the nb x nb block at A (lower part), identity-padded to 64 x 64, zeros above the diagonal (all 16 loads of a thread in flight)
__device__ __forceinline__ void load_diag_smem(double* __restrict__ T, const double* __restrict__ A, int lda, int nb) {
    double v[TB * TB / CT];
#pragma unroll
    for (int q = 0; q < TB * TB / CT; ++q) {
        int r, c; blk_rc(threadIdx.x + q * CT, r, c);
        v[q] = (r == c) ? 1.0 : 0.0;
        if (r < nb && c < nb) v[q] = (c <= r) ? A[(size_t)r + (size_t)c * lda] : 0.0;
    }
#pragma unroll
    for (int q = 0; q < TB * TB / CT; ++q) {
        int r, c; blk_rc(threadIdx.x + q * CT, r, c);
        T[r * LDT + c] = v[q];
    }
}

// T <- A_diag - P P' for the next diagonal block: A_diag (nb x nb, lower part) and the panel block P = L(k+1, k) (nb x 64) both straight from
// global memory into shared memory (all 32 loads of a thread in flight), the product on the shared-memory copy of P -- which is both
// operands -- for the three 32 x 32 blocks of the lower triangle only.  T comes out as load_diag_smem leaves it (identity-padded, zeros above).
__device__ __forceinline__ void update_diag_smem(double* __restrict__ T, double* __restrict__ Pm, const double* __restrict__ Adiag,
                                                 const double* __restrict__ Ppanel, int lda, int nb) {
    double v[TB * TB / CT], q_[TB * TB / CT];
#pragma unroll
    for (int q = 0; q < TB * TB / CT; ++q) {
        int r, c; blk_rc(threadIdx.x + q * CT, r, c);
        v[q] = (r == c) ? 1.0 : 0.0;
        if (r < nb && c < nb) v[q] = (c <= r) ? Adiag[(size_t)r + (size_t)c * lda] : 0.0;
        q_[q] = r < nb ? Ppanel[(size_t)r + (size_t)c * lda] : 0.0;
    }
#pragma unroll
    for (int q = 0; q < TB * TB / CT; ++q) {
        int r, c; blk_rc(threadIdx.x + q * CT, r, c);
        T[r * LDT + c] = v[q];
        Pm[r * LDT + c] = q_[q];
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int wr = warp >> 1, wc = warp & 1;
#pragma unroll
    for (int blk = 0; blk < 3; ++blk) {
        const int bi = blk == 0 ? 0 : 1, bj = blk == 2 ? 1 : 0;          // (0,0), (1,0), (1,1)
        const double* Pa = Pm + 32 * bi * LDT; const double* Pb = Pm + 32 * bj * LDT;
        const Frag32 f0 = smem_mma32(Pa, LDT, 1, Pb, 1, LDT);
        const Frag32 f1 = smem_mma32(Pa + 32, LDT, 1, Pb + 32, 1, LDT);
        double* C = T + 32 * bi * LDT + 32 * bj;
#pragma unroll
        for (int jj = 0; jj < 2; ++jj)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int r = wr * 8 + (lane >> 2), c = wc * 16 + 8 * jj + 2 * (lane & 3) + h;
                if (bi != bj || c <= r) C[r * LDT + c] -= f0.v[jj][h] + f1.v[jj][h];
            }
    }
    __syncthreads();
}

struct DenseJob {
    int M;
    int build;                 // 0: A as given | 1: A = P + w S2, xi = xip + w s1 (carry: P <- A, xip <- xi) | 2: A = K_uu(Z) + jitter I | 3: A = Sig + mu mu'
    double* A;                 // M x M column-major, factored in place: L in the lower triangle, strict upper triangle zeroed
    double* Dinv;              // inverses of the 64 x 64 diagonal blocks of L
    int* info;                 // 0, or 1 + the row of the first non-positive pivot
    const double* S2; const double* s1; double* P; double* xip; double* xi; double w; int carry;     // build 1
    const double* Z; int D, kind; double variance, jitter; double ell_inv[SGP_MAX_D];               // build 2
    const double* Sig; const double* mu_in;                                                         // build 3
    double* X; double* Tmp; double* S;     // optional: X = L^-1 (lower), S = X' X = (L L')^-1 full symmetric; Tmp = M x M scratch
    double* mu;                            // optional (needs S and xi): mu = S xi
    double* Ut;                            // optional: Ut = L' (upper triangular, strict lower part zero)
    int flip;                              // build 1 only: factorise J A J (J = index reversal); S is then scratch and Sout / mu / Uv are delivered un-reversed
    double* Sout;                          // flip: A^-1 (full symmetric)
    double* Uv;                            // flip, optional (needs mu, M <= kFlipUvMaxM): upper Cholesky factor of A^-1 + mu mu'
    long long* clk;                        // optional: CTA 0's clocks {build, factor, panel, trailing (+ X rows / S updates on the last step), barriers, mu, last S update, tail | inside factor: ...}
};

// One or TWO independent jobs of the same M per launch: every job is a serial chain on its own CTA 0 with the other CTAs waiting on it most of
// the time, so a second job (the N-th prod's Lambda and K_uu at a new theta, in the mini-batch schedule) runs in the shadow of the first.  CTAs are
// dealt alternately; both halves execute the same grid barriers (same block count; `sync_build` / `sync_tail` are the OR over the jobs).
struct DenseJobs {
    DenseJob job[2];
    int njobs, sync_build, sync_tail;
};

__global__ void __launch_bounds__(CT, 1) dense_job_kernel(const __grid_constant__ DenseJobs jj) {
    const int jsel = jj.njobs == 2 ? (int)(blockIdx.x & 1u) : 0;
    const DenseJob& j = jj.job[jsel];
    extern __shared__ double sm[];
    double* T = sm; double* Xi = T + TB * LDT; double* rdiag = Xi + TB * LDT; double* As = rdiag + TB + 2; double* Bs = As + STAGE_A;
    double* Lc = As;        // the factorisation's column-major scratch (32 x LCS) shares the GEMM staging area: never live together
    cg::grid_group grid = cg::this_grid();
    const int ncta = jj.njobs == 2 ? ((int)gridDim.x + 1 - jsel) / 2 : (int)gridDim.x, cta = jj.njobs == 2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
    const int M = j.M, nblk = (M + TB - 1) / TB, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const size_t MM = (size_t)M * M;
    double* __restrict__ A = j.A;
    long long tc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, fc[4] = {0, 0, 0, 0}, t0 = clock64(), t1;
    long long* fst = (j.clk && threadIdx.x == 0) ? fc : nullptr;
    long long tkl[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    long long* tk32 = j.clk ? tkl : nullptr;          // thread 0 of CTA 0 reports {barrier, load+stage, mma, tasks, chunks} of its 32 x 32 tasks
    long long* tk64 = j.clk ? tkl + 8 : nullptr;      // ... and of its 64 x 64 diagonal-block updates
#define DCLK(i) do { if (j.clk) { t1 = clock64(); tc[i] += t1 - t0; t0 = t1; } } while (0)

    // ---- build -------------------------------------------------------------------------------------------------------------------
    if (j.build == 1) {
        const size_t stride = (size_t)ncta * CT;
        for (size_t e0 = (size_t)cta * CT + tid; e0 < MM; e0 += 4 * stride) {        // (four independent loads in flight per thread)
            double s2[4], pr[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const size_t e = e0 + u * stride, src = j.flip ? MM - 1 - e : e;     // (reversing rows and columns reverses the column-major index)
                if (e < MM) { s2[u] = j.S2[src]; pr[u] = j.P[src]; }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const size_t e = e0 + u * stride, src = j.flip ? MM - 1 - e : e;
                if (e < MM) {
                    const double a = fma(j.w, s2[u], pr[u]);
                    A[e] = a;
                    if (j.carry) j.P[src] = a;
                }
            }
        }
        if (cta == ncta - 1)
            for (int e = tid; e < M; e += CT) {
                const double x = fma(j.w, j.s1[e], j.xip[e]);
                j.xi[e] = x;
                if (j.carry) j.xip[e] = x;
            }
    } else if (j.build == 2) {
#pragma unroll 2
        for (size_t e = (size_t)cta * CT + tid; e < MM; e += (size_t)ncta * CT) {
            const int r = (int)(e % M), c = (int)(e / M);
            double r2 = 0.0;
            for (int d = 0; d < j.D; ++d) { const double t = (j.Z[(size_t)r * j.D + d] - j.Z[(size_t)c * j.D + d]) * j.ell_inv[d]; r2 = fma(t, t, r2); }
            double v;
            if (j.kind == SGP_KERNEL_SE) v = j.variance * exp(-0.5 * r2);
            else if (j.kind == SGP_KERNEL_MATERN32) { const double s = sqrt(3.0 * r2); v = j.variance * (1.0 + s) * exp(-s); }
            else { const double s = sqrt(5.0 * r2); v = j.variance * (1.0 + s + s * s / 3.0) * exp(-s); }
            if (r == c) v += j.jitter;
            A[e] = v;
        }
    } else if (j.build == 3) {
        for (size_t e = (size_t)cta * CT + tid; e < MM; e += (size_t)ncta * CT) A[e] = fma(j.mu_in[e % M], j.mu_in[e / M], j.Sig[e]);
    }
    if (jj.sync_build) grid.sync();  // (grid.sync orders every thread's earlier stores: no explicit fence in front of the barriers)
    DCLK(0);

    // ---- Cholesky ----------------------------------------------------------------------------------------------------------------
    if (cta == 0) {
        load_diag_smem(T, A, M, min(TB, M));
        factor_diag_smem(T, Xi, rdiag, Lc, A, M, min(TB, M), 0, j.Dinv, j.info, fst);
    }
    DCLK(1);
    grid.sync();
    DCLK(4);
    // The inverse rides along: while CTA 0 factorises diagonal block k + 1 (the critical path, ~24 k clocks), the other CTAs -- after the trailing
    // update of step k -- build ROW BLOCK k of X = L^-1 and add row block k - 1 of X to S = X'X.  Right-looking, so that every task is short
    // (K = 64): with P(i, j) = sum_{l <= k-2} L(i, l) X(l, j) kept up to date in `Tmp` for the rows below,
    //     X(k, j) = -Dinv_k [ P(k, j) + L(k, k-1) X(k-1, j) ],   X(k, k) = Dinv_k,     then   P(i, j) += L(i, k-1) X(k-1, j)  for i > k.
    // Everything read in a step was finished in an earlier one: no separate inverse phases, no extra grid barriers, and the workers stay inside
    // CTA 0's factorisation window.
    double* __restrict__ X = j.X;
    auto x_slice = [&](int k, int t) {                        // slice t of row block k of X (64 x 16 columns; t == 4 k: the diagonal block)
        const int kb = k * TB, nbk = min(TB, M - kb);
        const double* Dk = j.Dinv + (size_t)k * TB * TB;
        if (t == 4 * k) {                                      // X(k, k) = Dinv_k
            double v[TB * TB / CT];
#pragma unroll
            for (int q = 0; q < TB * TB / CT; ++q) v[q] = Dk[tid + q * CT];
#pragma unroll
            for (int q = 0; q < TB * TB / CT; ++q) {
                const int e = tid + q * CT, r = e % TB, c = e / TB;
                if (r < nbk && c < nbk) X[(size_t)(kb + r) + (size_t)(kb + c) * M] = v[q];
            }
            return;
        }
        const int c0 = 16 * t;
        double* Tt = j.Tmp + (size_t)kb + (size_t)c0 * M;      // P(k, slice), completed in place with the l = k - 1 term
        TX::Acc acc; acc_zero<TX>(acc);
        gemm_task<TX>(acc, A + (size_t)kb + (size_t)(kb - TB) * M, 1, (size_t)M, nbk, X + (size_t)(kb - TB) + (size_t)c0 * M, 1, (size_t)M, 16, TB, As, Bs);
        store_task<TX>(acc, Tt, M, nbk, 16, 1.0, (c0 / TB == k - 1) ? 0.0 : 1.0);
        __syncthreads();                                       // (the block reads back its own global stores)
        TX::Acc acc2; acc_zero<TX>(acc2);
        gemm_task<TX>(acc2, Dk, 1, (size_t)TB, nbk, Tt, 1, (size_t)M, 16, nbk, As, Bs);
        store_task<TX>(acc2, X + (size_t)kb + (size_t)c0 * M, M, nbk, 16, -1.0, 0.0);
    };
    auto tri_ab = [](int t, int& a, int& b) { a = 0; while ((a + 1) * (a + 2) / 2 <= t) ++a; b = t - a * (a + 1) / 2; };
    auto trailing_tile = [&](int k, int t) {                  // 64 x 64 tile t + 1 of the lower triangle of A[R0:, R0:] (tile 0 belongs to CTA 0)
        const int k0 = k * TB, R0 = k0 + TB;
        int a, b; tri_ab(t + 1, a, b);
        const int r0 = R0 + TB * a, c0 = R0 + TB * b;
        const Task32 tk{A + (size_t)r0 + (size_t)k0 * M, 1, (size_t)M, min(TB, M - r0), A + (size_t)c0 + (size_t)k0 * M, (size_t)M, 1, min(TB, M - c0), TB,
                        A + (size_t)r0 + (size_t)c0 * M, M, -1.0, 1.0, nullptr, a == b, r0, c0, tk32};
        run_task64(tk, As, Bs);
    };
    auto p_tile = [&](int k, int t) {                          // P(i, jb) (+)= L(i, k-1) X(k-1, jb), row blocks i > k, column blocks jb < k
        const int nr = nblk - k - 1, kb1 = (k - 1) * TB;
        const int r0 = (k + 1 + t % nr) * TB, jb = t / nr, c0 = jb * TB;
        const Task32 tk{A + (size_t)r0 + (size_t)kb1 * M, 1, (size_t)M, min(TB, M - r0), X + (size_t)kb1 + (size_t)c0 * M, 1, (size_t)M, TB, TB,
                        j.Tmp + (size_t)r0 + (size_t)c0 * M, M, 1.0, jb == k - 1 ? 0.0 : 1.0, nullptr, false, 0, 0, tk32};
        run_task64(tk, As, Bs);
    };
    auto s_tile = [&](int kk, int t, bool last) {              // S(a, b) (+)= X(kk, a)' X(kk, b), blocks b <= a <= kk
        const int kb = kk * TB, nbk = min(TB, M - kb);
        int a, b; tri_ab(t, a, b);
        const int r0 = TB * a, c0 = TB * b;
        const Task32 tk{X + (size_t)kb + (size_t)r0 * M, (size_t)M, 1, min(TB, M - r0), X + (size_t)kb + (size_t)c0 * M, 1, (size_t)M, min(TB, M - c0), nbk,
                        j.S + (size_t)r0 + (size_t)c0 * M, M, 1.0, a == kk ? 0.0 : 1.0, last ? j.S + (size_t)c0 + (size_t)r0 * M : nullptr, a == b, r0, c0, tk32};
        run_task64(tk, As, Bs);
    };
    auto s_tile32 = [&](int kk, int t, bool last) {            // the same in 32 x 32 sub-tiles: for the phases everybody waits for, when they are small
        const int kb = kk * TB, nbk = min(TB, M - kb);
        int a, b; tri_ab(t, a, b);
        const int r0 = 32 * a, c0 = 32 * b;
        const Task32 tk{X + (size_t)kb + (size_t)r0 * M, (size_t)M, 1, min(32, M - r0), X + (size_t)kb + (size_t)c0 * M, 1, (size_t)M, min(32, M - c0), nbk,
                        j.S + (size_t)r0 + (size_t)c0 * M, M, 1.0, r0 / TB == kk ? 0.0 : 1.0, last ? j.S + (size_t)c0 + (size_t)r0 * M : nullptr, a == b, r0, c0, tk32};
        run_task32(tk, As, Bs);
    };
    // All of a step's worker tasks in ONE list dealt round-robin (tiles first, the lighter X slices last): nobody reads what another task of the
    // same step writes.
    auto step_tasks = [&](int k, int w, int nw) {
        const bool has_next = k + 1 < nblk;
        const int nb = has_next ? nblk - k - 1 : 0;
        const int n_tr = has_next ? nb * (nb + 1) / 2 - 1 : 0;
        const int n_p = (X && k >= 1) ? nb * k : 0;
        int n_s = (X && j.S && k >= 1) ? k * (k + 1) / 2 : 0;
        const int n_x = X ? 4 * k + 1 : 0;
        const bool s32 = !has_next && n_s > 0 && 3 * n_s <= kFineTiles;           // last step (nobody factorises meanwhile): finer tasks when they fit one round
        if (s32) n_s = k * (2 * k + 1);
        for (int t = w; t < n_tr + n_p + n_s + n_x; t += nw) {
            if (t < n_tr) trailing_tile(k, t);
            else if (t < n_tr + n_p) p_tile(k, t - n_tr);
            else if (t < n_tr + n_p + n_s) { if (s32) s_tile32(k - 1, t - n_tr - n_p, false); else s_tile(k - 1, t - n_tr - n_p, false); }
            else x_slice(k, t - n_tr - n_p - n_s);
        }
    };
    for (int k = 0; k < nblk; ++k) {
        const bool has_next = k + 1 < nblk;
        const int k0 = k * TB, R0 = k0 + TB;                   // trailing matrix starts at row / column R0
        const double* Dk = j.Dinv + (size_t)k * TB * TB;
        if (has_next) {
            // panel: 16-row strips of A[R0:, k0:k0+64] <- strip * Dk'   (in place: a strip is private to its CTA)
            const int nstrips = (M - R0 + 15) / 16;
            for (int s = cta; s < nstrips; s += ncta) {
                const int r0 = R0 + 16 * s, rows = min(16, M - r0);
                double* Ar = A + (size_t)r0 + (size_t)k0 * M;
                T16::Acc acc; acc_zero<T16>(acc);
                gemm_task<T16>(acc, Ar, 1, (size_t)M, rows, Dk, (size_t)TB, 1, TB, TB, As, Bs);      // B(kk, c) = Dk(c, kk)
                store_task<T16>(acc, Ar, M, rows, TB, 1.0, 0.0);
            }
            DCLK(2);
            grid.sync();
            DCLK(4);
        }
        // CTA 0 factorises the next diagonal block; everybody else: trailing update, row block k of X, row block k - 1 of X into S
        const bool factoring = has_next && cta == 0 && ncta > 1;
        const int nwork = (has_next && ncta > 1) ? ncta - 1 : ncta, wid = (has_next && ncta > 1) ? cta - 1 : cta;
        if (has_next && cta == 0) {     // next diagonal block: A[i, c] -= sum_kk L[i, k0 + kk] L[c, k0 + kk], straight into shared memory, factorised there
            const int nbn = min(TB, M - R0);
            update_diag_smem(T, Xi, A + (size_t)R0 * ((size_t)M + 1), A + (size_t)R0 + (size_t)k0 * M, M, nbn);
            DCLK(3);
            factor_diag_smem(T, Xi, rdiag, Lc, A + (size_t)R0 * ((size_t)M + 1), M, nbn, R0, j.Dinv + (size_t)(k + 1) * TB * TB, j.info, fst);
            DCLK(1);
        }
        if (!factoring) step_tasks(k, wid, nwork);
        DCLK(3);
        grid.sync();
        DCLK(4);
    }
    if (X && j.S) {       // the last row block of X into S (every tile receives its last contribution here: mirrored on the way out), then mu
        if (3 * nblk * (nblk + 1) / 2 <= kFineTiles) {
            const int n32 = (M + 31) / 32;
            for (int t = cta; t < n32 * (n32 + 1) / 2; t += ncta) s_tile32(nblk - 1, t, true);
        } else {
            for (int t = cta; t < nblk * (nblk + 1) / 2; t += ncta) s_tile(nblk - 1, t, true);
        }
        if (j.flip && j.mu) {
            // p = X xi (reversed indices) in kPParts partial sums per 32-row strip: rows = lanes, a warp per column, fixed-order reduction.
            // Needs X only: dealt from the far end of the grid, beside the S tiles.
            const int nstrip = (M + 31) / 32;
            double* red = sm;                                  // 8 warps x 32 rows
            for (int t = ncta - 1 - cta; t < nstrip * kPParts; t += ncta) {
                const int st = t / kPParts, q = t % kPParts, row = 32 * st + lane;
                const int ncol = min(M, 32 * st + 32), cb = (int)((long long)ncol * q / kPParts), ce = (int)((long long)ncol * (q + 1) / kPParts);
                double v = 0.0;
                if (row < M)
#pragma unroll 4
                    for (int c = cb + warp; c < ce; c += CT / 32)
                        if (c <= row) v = fma(X[(size_t)row + (size_t)c * M], j.xi[M - 1 - c], v);
                __syncthreads();
                red[warp * 32 + lane] = v;
                __syncthreads();
                if (warp == 0 && row < M) {
                    double a = red[lane];
#pragma unroll
                    for (int w8 = 1; w8 < CT / 32; ++w8) a += red[w8 * 32 + lane];
                    j.Tmp[(size_t)q * M + row] = a;
                }
            }
        }
        DCLK(6);
    }
    if (jj.sync_tail) grid.sync();
    if (X && j.S) {
        if (!j.flip) {
            if (j.mu)          // mu = S xi: a warp per column of the symmetric S (fixed-shape tree: deterministic)
                for (int i = cta * (CT / 32) + warp; i < M; i += ncta * (CT / 32)) {
                    const double* col = j.S + (size_t)i * M;
                    double v = 0.0;
                    for (int r = lane; r < M; r += 32) v = fma(col[r], j.xi[r], v);
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                    if (lane == 0) j.mu[i] = v;
                }
        } else {
            // A^-1 = U0' U0 with U0 = J X J upper triangular: the Cholesky factor of the inverse is the inverse of the REVERSED-order factor.
            //   mu = U0' p, p = U0 xi (two triangular products, as a Cholesky solve would do), and
            //   A^-1 + mu mu' = U0' (I + p p') U0, where the factor G of I + p p' is closed-form:
            //     G[k, k] = s_{k+1} / s_k,   G[k, i] = p_k p_i / (s_k s_{k+1})  (i > k),   s_k^2 = 1 + sum_{i < k} p_i^2,
            // so Uv = G U0 costs a running sum down every column of X instead of a second M^3 / 3 factorisation:
            //     Uv[k, c] = (s_{k+1} / s_k) U0[k, c] + p_k / (s_k s_{k+1}) sum_{k < i <= c} p_i U0[i, c].
            double* pp = sm; double* ga = sm + M; double* b0 = sm + 2 * M; double* b1 = sm + 3 * M;     // (in reversed indices throughout)
            const double* gb = nullptr;
            if (j.mu) {
                for (int i = tid; i < M; i += CT) {
                    double v = j.Tmp[i];
#pragma unroll
                    for (int q = 1; q < kPParts; ++q) v += j.Tmp[(size_t)q * M + i];
                    pp[i] = v; b0[i] = v * v;
                }
                __syncthreads();
            }
            if (j.Uv) {
                double* cur = b0; double* nxt = b1;                 // inclusive suffix sums of p^2 (positive terms: no cancellation)
                for (int off = 1; off < M; off <<= 1) {
                    for (int i = tid; i < M; i += CT) nxt[i] = cur[i] + (i + off < M ? cur[i + off] : 0.0);
                    __syncthreads();
                    double* sw = cur; cur = nxt; nxt = sw;
                }
                for (int i = tid; i < M; i += CT) {                 // s_k^2 = 1 + (sum over reversed indices above i), s_{k+1}^2 = s_k^2 + p_i^2
                    const double lo = 1.0 + (i + 1 < M ? cur[i + 1] : 0.0), hi = 1.0 + cur[i];
                    ga[i] = sqrt(hi / lo);
                    nxt[i] = pp[i] / sqrt(hi * lo);
                }
                __syncthreads();
                gb = nxt;
            }
            // a warp per column c of X (column M - 1 - c of Uv), from the diagonal down (Uv: from the diagonal up), 128 rows at a time:
            // mu[M - 1 - c] = sum_r X[r, c] p[r] and the column of Uv in the same pass
            if (j.mu)
                for (int c = cta + ncta * warp; c < M; c += ncta * (CT / 32)) {
                    const double* xc = X + (size_t)c * M;
                    double* uc = j.Uv ? j.Uv + (size_t)(M - 1 - c) * M : nullptr;
                    if (uc) for (int r = lane; r < c; r += 32) uc[M - 1 - r] = 0.0;       // strictly below the diagonal of Uv
                    double carry = 0.0, dot = 0.0;
                    for (int r0 = c; r0 < M; r0 += 128) {
                        double x[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) { const int r = r0 + 32 * u + lane; x[u] = r < M ? xc[r] : 0.0; }
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const int r = r0 + 32 * u + lane;
                            const double t = r < M ? pp[r] * x[u] : 0.0;
                            dot += t;
                            if (uc) {
                                double inc = t;
#pragma unroll
                                for (int o = 1; o < 32; o <<= 1) { const double y = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += y; }
                                if (r < M) uc[M - 1 - r] = fma(gb[r], carry + (inc - t), ga[r] * x[u]);
                                carry += __shfl_sync(0xffffffffu, inc, 31);
                            }
                        }
                    }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
                    if (lane == 0) j.mu[M - 1 - c] = dot;
                }
            {
                const size_t stride = (size_t)ncta * CT;
                for (size_t e0 = (size_t)cta * CT + tid; e0 < MM; e0 += 8 * stride) {
                    double v[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) { const size_t e = e0 + u * stride; if (e < MM) v[u] = j.S[MM - 1 - e]; }
#pragma unroll
                    for (int u = 0; u < 8; ++u) { const size_t e = e0 + u * stride; if (e < MM) j.Sout[e] = v[u]; }
                }
            }
        }
        DCLK(5);
    }

    // ---- tail: zero the strict upper triangle of L; Ut = L' -----------------------------------------------------------------------------
    if (j.Ut) {
        const int n32 = (M + 31) / 32;
        double* tile = sm;                                         // 32 x 33
        for (int t = cta; t < n32 * n32; t += ncta) {
            const int bi = t % n32, bj = t / n32;                  // block (bi, bj) of L -> block (bj, bi) of Ut
            __syncthreads();
            for (int e = tid; e < 32 * 32; e += CT) {
                const int r = bi * 32 + e % 32, c = bj * 32 + e / 32;
                double v = 0.0;
                if (r < M && c < M && c <= r) v = A[(size_t)r + (size_t)c * M];
                tile[(e % 32) * 33 + e / 32] = v;
            }
            __syncthreads();
            for (int e = tid; e < 32 * 32; e += CT) {
                const int c = bj * 32 + e % 32, r = bi * 32 + e / 32;          // Ut[c, r] = L[r, c]
                if (r < M && c < M) {
                    j.Ut[(size_t)c + (size_t)r * M] = tile[(e / 32) * 33 + e % 32];
                    if (c > r) A[(size_t)r + (size_t)c * M] = 0.0;
                }
            }
        }
    } else if (!j.flip) {                                              // (nobody reads the factor of a reversed-order job)
        for (size_t e = (size_t)cta * CT + tid; e < MM; e += (size_t)ncta * CT)
            if (e / M > e % M) A[e] = 0.0;
    }
    DCLK(7);
    if (j.clk && cta == 0 && tid == 0) {
        for (int i = 0; i < 8; ++i) j.clk[i] = tc[i];
        for (int i = 0; i < 4; ++i) j.clk[8 + i] = fc[i];
        for (int i = 0; i < 8; ++i) { j.clk[12 + i] = tkl[i]; j.clk[20 + i] = tkl[8 + i]; }
    }
#undef DCLK
}

// General GEMM on the same tile routine: C[m x n] = beta C + alpha op(A) op(B); 32 x 32 tiles while they fit one wave, else 64 x 64.
struct Gemm2Args {
    const double* A; const double* B; double* C;
    int m, n, k, lda, ldb, ldc, opA, opB, lower_only;
    double alpha, beta;
};
template <class TL, int TSR, int TSC>
__global__ void __launch_bounds__(CT) gemm2_kernel(const Gemm2Args g) {
    extern __shared__ double sm[];
    double* As = sm; double* Bs = sm + STAGE_A;
    const int tm = (g.m + TSR - 1) / TSR, tn = (g.n + TSC - 1) / TSC;
    for (int t = blockIdx.x; t < tm * tn; t += gridDim.x) {
        const int ti = t % tm, tj = t / tm;
        if (g.lower_only && tj * TSC > ti * TSR + TSR - 1) continue;      // tile entirely above the diagonal
        const int rows = min(TSR, g.m - ti * TSR), cols = min(TSC, g.n - tj * TSC);
        // A(r, k): opA == 0 -> A[(ti*TSR + r) + k*lda], else A[k + (ti*TSR + r)*lda];  B(k, c): opB == 0 -> B[k + (tj*TSC + c)*ldb], else B[(tj*TSC + c) + k*ldb]
        const double* Ab = g.opA == 0 ? g.A + (size_t)ti * TSR : g.A + (size_t)ti * TSR * g.lda;
        const double* Bb = g.opB == 0 ? g.B + (size_t)tj * TSC * g.ldb : g.B + (size_t)tj * TSC;
        typename TL::Acc acc; acc_zero<TL>(acc);
        const int kk = g.lower_only == 2 ? min(g.k, (tj + 1) * TSC) : g.k;      // C = U'U, U upper triangular: column c needs k <= c only
        if constexpr (TL::MI * TL::NJ > 2)
            gemm_task_db<TL>(acc, Ab, g.opA == 0 ? 1 : (size_t)g.lda, g.opA == 0 ? (size_t)g.lda : 1, rows, Bb, g.opB == 0 ? 1 : (size_t)g.ldb,
                             g.opB == 0 ? (size_t)g.ldb : 1, cols, kk, sm);
        else
            gemm_task<TL>(acc, Ab, g.opA == 0 ? 1 : (size_t)g.lda, g.opA == 0 ? (size_t)g.lda : 1, rows, Bb, g.opB == 0 ? 1 : (size_t)g.ldb,
                          g.opB == 0 ? (size_t)g.ldb : 1, cols, kk, As, Bs);
        store_task<TL>(acc, g.C + (size_t)ti * TSR + (size_t)tj * TSC * g.ldc, g.ldc, rows, cols, g.alpha, g.beta);
    }
}
template <class TL, int TSR, int TSC>
int launch_gemm2(sgp_ctx* ctx, const Gemm2Args& g, int grid) {
    const size_t smem = TL::MI * TL::NJ > 2 ? (size_t)2 * (TSR * TL::LDA + TL::KC * TL::LDB) * sizeof(double) : (size_t)(STAGE_A + STAGE_B) * sizeof(double);
    SGP_CUDA(ctx, cudaFuncSetAttribute(gemm2_kernel<TL, TSR, TSC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    gemm2_kernel<TL, TSR, TSC><<<grid, CT, smem, ctx->stream>>>(g);
    SGP_CUDA(ctx, cudaGetLastError());
    return SGP_OK;
}

}  // namespace

// One cooperative launch of dense_job_kernel on the ctx stream: job `in`, and optionally a second independent job `in2` of the same M beside
// it (pivot record of the second job: ctx->info_dev[1]).  The caller checks *info_dev (sgp_dense_info) when it next synchronises.
static int fill_job(sgp_ctx* ctx, const SgpDenseJob& in, DenseJob& j, int* info) {
    j.M = in.M; j.build = in.build; j.A = in.A; j.Dinv = in.Dinv; j.info = info;
    j.S2 = in.S2; j.s1 = in.s1; j.P = in.P; j.xip = in.xip; j.xi = in.xi; j.w = in.w; j.carry = in.carry;
    j.Z = ctx->Z_dev; j.D = ctx->D; j.kind = ctx->kind; j.variance = ctx->variance; j.jitter = in.jitter;
    for (int d = 0; d < SGP_MAX_D; ++d) j.ell_inv[d] = d < ctx->D ? 1.0 / ctx->ell[d] : 0.0;
    j.Sig = in.Sig; j.mu_in = in.mu_in; j.X = in.X; j.Tmp = in.Tmp; j.S = in.S; j.mu = in.mu; j.Ut = in.Ut; j.clk = in.clk;
    j.flip = in.flip; j.Sout = in.Sout; j.Uv = in.Uv;
    if (j.mu && !(j.S && j.xi)) SGP_FAIL(ctx, SGP_ERR_ARG, "dense job: mu needs S and xi");
    if (j.flip && !(j.build == 1 && j.X && j.S && j.Sout)) SGP_FAIL(ctx, SGP_ERR_ARG, "dense job: the reversed order is for build 1 with X, S and Sout");
    if (j.Uv && !(j.flip && j.mu && j.M <= SGP_FLIP_UV_MAX_M)) SGP_FAIL(ctx, SGP_ERR_ARG, "dense job: Uv needs the reversed order, mu and M <= 4096");
    return SGP_OK;
}
int sgp_dense_job(sgp_ctx* ctx, const SgpDenseJob& in, const SgpDenseJob* in2) {
    DenseJobs jj{};
    int rc = fill_job(ctx, in, jj.job[0], ctx->info_dev); if (rc) return rc;
    jj.njobs = 1;
    if (in2) {
        if (in2->M != in.M) SGP_FAIL(ctx, SGP_ERR_ARG, "dense job: two jobs in one launch need the same M");
        rc = fill_job(ctx, *in2, jj.job[1], ctx->info_dev + 1); if (rc) return rc;
        jj.njobs = 2;
    }
    for (int q = 0; q < jj.njobs; ++q) {
        const DenseJob& j = jj.job[q];
        jj.sync_build |= j.build != 0;
        jj.sync_tail |= (j.X && j.S && (j.mu || j.flip)) ? 1 : 0;
    }
    DenseJob& j = jj.job[0];
    const int M = in.M, nblk = (M + TB - 1) / TB, n32 = (M + 31) / 32;
    const size_t smem = SMEM_DOUBLES * sizeof(double);
    SGP_CUDA(ctx, cudaFuncSetAttribute(dense_job_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    static const bool print_clocks = std::getenv("SGP_DENSE_CLOCKS") != nullptr;       // tuning aid: per-phase clocks of CTA 0 on stdout
    long long* clk_dev = nullptr;
    if (print_clocks && !j.clk) { SGP_CUDA(ctx, cudaMalloc((void**)&clk_dev, 28 * sizeof(long long))); SGP_CUDA(ctx, cudaMemsetAsync(clk_dev, 0, 28 * sizeof(long long), ctx->stream)); j.clk = clk_dev; }
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, dense_job_kernel, CT, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
    int want = std::max(1, n32 * (n32 + 1) / 2);                    // the widest phases: trailing sub-tiles of the first panel, S = X'X
    if (nblk == 1 && !in.X) want = 1;
    if (in2) want = std::max(2, 2 * want);
    const int grid = std::max(jj.njobs, std::min(want, per_sm * ctx->num_sms));
    if (in.reset_info) SGP_CUDA(ctx, cudaMemsetAsync(ctx->info_dev, 0, 2 * sizeof(int), ctx->stream));
    void* args[] = {&jj};
    SGP_CUDA(ctx, cudaLaunchCooperativeKernel((const void*)dense_job_kernel, dim3(grid), dim3(CT), args, smem, ctx->stream));
    SGP_CUDA(ctx, cudaGetLastError());
    if (clk_dev) {
        long long c[28];
        SGP_CUDA(ctx, cudaMemcpyAsync(c, clk_dev, sizeof c, cudaMemcpyDeviceToHost, ctx->stream));
        SGP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        cudaFree(clk_dev);
        printf("dense job M=%d build=%d jobs=%d grid=%d: clocks build %lld | factor %lld | panel %lld | next diagonal block (+ the last step's tasks) %lld | barriers %lld | mu, Uv %lld | last S update %lld | tail %lld || factor: chol32 %lld | inv32 %lld | products %lld | write-out %lld\n",
               M, in.build, jj.njobs, grid, c[0], c[1], c[2], c[3], c[4], c[5], c[6], c[7], c[8], c[9], c[10], c[11]);
        printf("    CTA 0's tile tasks: %lld tasks, %lld chunks: barrier %lld | load+stage %lld | mma %lld\n", c[15], c[16], c[12], c[13], c[14]);
        fflush(stdout);
    }
    return SGP_OK;
}

// Synchronises the stream and turns a recorded non-positive pivot into SGP_ERR_NOT_PD.
int sgp_dense_info(sgp_ctx* ctx, const char* what) {
    int both[2] = {0, 0};
    SGP_CUDA(ctx, cudaMemcpyAsync(both, ctx->info_dev, 2 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    SGP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    const int info = both[0];
    ctx->info2_last = both[1];             // (second job of a two-job launch)
    if (info != 0) {
        char buf[160];
        snprintf(buf, sizeof buf, "%s: Cholesky met a non-positive pivot at row %d of %d", what, info, ctx->M);
        SGP_FAIL(ctx, SGP_ERR_NOT_PD, buf);
    }
    return SGP_OK;
}

// In-place lower Cholesky (column-major, upper triangle zeroed); the inverses of the 64 x 64 diagonal blocks are left in
// ctx->dinv_dev.  Non-positive pivot -> SGP_ERR_NOT_PD.
int sgp_potrf_lower(sgp_ctx* ctx, double* A, int M) {
    const int nblk = (M + TB - 1) / TB;
    int rc = sgp_ensure_zero(ctx, &ctx->dinv_dev, &ctx->dinv_cap, (size_t)nblk * TB * TB); if (rc) return rc;
    SgpDenseJob j{};
    j.M = M; j.A = A; j.Dinv = ctx->dinv_dev; j.reset_info = 1;
    rc = sgp_dense_job(ctx, j); if (rc) return rc;
    return sgp_dense_info(ctx, "potrf");
}

int sgp_gemm2(sgp_ctx* ctx, int opA, int opB, int m, int n, int k, double alpha, const double* A, int lda, const double* B, int ldb, double beta,
              double* C, int ldc, int lower_only) {
    if (m <= 0 || n <= 0) return SGP_OK;
    Gemm2Args g{A, B, C, m, n, k, lda, ldb, ldc, opA, opB, lower_only, alpha, beta};
    // Tile shape by the fewest (waves over the SMs) x (measured relative cost of a tile: 32 x 32 = 1, 64 x 32 = 1.5, 64 x 64 = 2.6 -- the
    // larger tiles move less operand traffic per flop, the smaller ones fill the machine when the matrix is small)
    const int sms = ctx->num_sms;
    const int t32 = ((m + 31) / 32) * ((n + 31) / 32), t63 = ((m + 63) / 64) * ((n + 31) / 32), t64 = ((m + 63) / 64) * ((n + 63) / 64);
    const double c32 = 1.0 * ((t32 + sms - 1) / sms), c63 = 1.5 * ((t63 + sms - 1) / sms), c64 = 2.6 * ((t64 + sms - 1) / sms);
    if (c32 <= c63 && c32 <= c64) return launch_gemm2<T32, 32, 32>(ctx, g, t32);
    if (c63 <= c64) return launch_gemm2<T63, 64, 32>(ctx, g, t63);
    return launch_gemm2<T64, 64, 64>(ctx, g, std::min(t64, 4 * sms));
}

// B (M x nrhs, ld M) <- L^-1 B (trans = false) or L^-T B (trans = true) with the diagonal-block inverses `dinv` of L
// (as left by the Cholesky): per 64-wide panel one small GEMM with the inverse block and one GEMM update.
int sgp_trsm_lower_dinv(sgp_ctx* ctx, const double* L, const double* dinv, double* B, double* tmp /* 64 x nrhs */, int M, int nrhs, bool trans) {
    const int nblk = (M + TB - 1) / TB;
    for (int s = 0; s < nblk; ++s) {
        const int kb = trans ? nblk - 1 - s : s;
        const int k = kb * TB, nb = std::min(TB, M - k);
        const double* Dk = dinv + (size_t)kb * TB * TB;
        // X_k = Dk B_k (or Dk' B_k): into tmp, then back
        int rc = sgp_gemm(ctx, trans ? 1 : 0, 0, nb, nrhs, nb, 1.0, Dk, TB, B + k, M, 0.0, tmp, TB, 0); if (rc) return rc;
        SGP_CUDA(ctx, cudaMemcpy2DAsync(B + k, (size_t)M * sizeof(double), tmp, (size_t)TB * sizeof(double), (size_t)nb * sizeof(double), (size_t)nrhs,
                                        cudaMemcpyDeviceToDevice, ctx->stream));
        if (!trans) {
            const int rows = M - k - nb;
            if (rows > 0) { rc = sgp_gemm(ctx, 0, 0, rows, nrhs, nb, -1.0, L + (size_t)(k + nb) + (size_t)k * M, M, B + k, M, 1.0, B + k + nb, M, 0); if (rc) return rc; }
        } else if (k > 0) {
            rc = sgp_gemm(ctx, 1, 0, k, nrhs, nb, -1.0, L + (size_t)k, M, B + k, M, 1.0, B, M, 0); if (rc) return rc;
        }
    }
    return SGP_OK;
}
