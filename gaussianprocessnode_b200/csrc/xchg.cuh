// Sum of the per-rank statistics over NVLink / NVSwitch peer memory, as device code that runs in the TAIL of the kernel that produced them
// (sweep4_kernel.cuh) or as a small kernel of its own (xchg.cu) -- no host involvement, no NCCL launch.  The reference has no
// distributed code; this is the one collective of the N-sharded sweep (section 8e).
//
// Every rank owns a region [flags (256 B) | contribution (cap doubles) | result (cap doubles)] that all peers map through CUDA IPC
// (comm.cu).  TWO-SHOT all-reduce for R <= 8 ranks, on the PACKED statistics (lower triangle of Psi2 column by column, M (M + 1) / 2
// doubles, then Psi1 and the four scalars):
//   1. the rank writes its contribution into its OWN contribution buffer.
//   2. the LAST CTA to finish (ticket counter, no grid barrier) publishes flag A = epoch into every rank's flag array (st.release.sys
//      after a system fence); one warp per CTA polls the LOCAL flag array until all R ranks have published.
//   3. reduce-scatter: rank r adds elements [r, r + 1) * n / R of all R contributions (16-byte ld.cv over NVLink, all ranks' loads of a
//      thread in flight together) in rank order -- every element is summed exactly once, so all ranks end up with identical bits --
//      and stores the sums into the result buffer of EVERY rank (16-byte stores over NVLink).
//   4. the last CTA publishes flag B = epoch everywhere; every CTA waits until all R shares have landed in the local result buffer.
//   5. every CTA expands its part of the local result into the full symmetric Psi2 | Psi1 | scalars of the statistics buffer.
// Measured on 8 B200 (tools/xchg_bench.py, kin40k shape): the one-shot forms move R times the volume per rank and cost 60 us (pull) /
// 78 us (push of 8-byte stores) at R = 8 against 29 / 42 us at R = 2.
// Flags are monotonic epochs, never reset; the ticket counters are reset by the CTA that completes them.  A rank may overwrite its
// contribution buffer in the next exchange without further handshake: it has seen flag B of every peer, and a peer signals B only after its
// reduce-scatter has finished reading.
#pragma once
#include "sgp_internal.cuh"

namespace sgp_xchg {

constexpr int kFlagA = 0, kFlagB = 16, kCntA = 32, kCntB = 33;     // word offsets inside the 256-byte flag block

__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) { asm volatile("st.release.sys.global.u32 [%0], %1;\n" ::"l"(p), "r"(v) : "memory"); }
// wait until *flag has reached `epoch` (a peer that never arrives means a lost rank: trap after ~20 s instead of hanging the GPU)
__device__ __forceinline__ void wait_epoch(const unsigned* flag, unsigned epoch) {
    for (int spin = 0; spin < 64; ++spin)
        if ((int)(ld_acquire_sys(flag) - epoch) >= 0) return;
    const long long t0 = clock64();
    while ((int)(ld_acquire_sys(flag) - epoch) < 0) {
        __nanosleep(40);
        if (clock64() - t0 > 40000000000ll) __trap();
    }
}
__device__ __forceinline__ unsigned* flags_of(const SgpXchg& x, int q) { return reinterpret_cast<unsigned*>(x.peers[q]); }
__device__ __forceinline__ double* contrib_of(const SgpXchg& x, int q) { return reinterpret_cast<double*>(x.peers[q] + x.slot0_off); }
__device__ __forceinline__ double* result_of(const SgpXchg& x, int q) { return reinterpret_cast<double*>(x.peers[q] + x.slot0_off + x.slot_bytes); }

// packed offset of column j of an M x M lower triangle stored column by column
__device__ __forceinline__ long long tri_col(long long j, long long M) { return j * M - j * (j - 1) / 2; }

// step 1: element e of this rank's contribution
__device__ __forceinline__ void put1(const SgpXchg& x, long long e, double v) { contrib_of(x, x.rank)[e] = v; }

// steps 2 / 4, first half: this CTA's stores are done; the last of the `ncta` CTAs publishes flag `which` (kFlagA / kFlagB) everywhere.
// The R flag stores go out in PARALLEL (thread q signals rank q after its own system fence): a release store per peer from one thread costs
// ~1.5 us EACH (measured: the barrier pair grew by ~3 us per rank).
__device__ __forceinline__ void publish(const SgpXchg& x, int ncta, int which) {
    __shared__ int s_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned* cnt = flags_of(x, x.rank) + (which == kFlagA ? kCntA : kCntB);
        const bool last = atomicAdd(cnt, 1u) == (unsigned)ncta - 1u;
        if (last) *cnt = 0u;
        s_last = last ? 1 : 0;
    }
    __syncthreads();
    if (s_last && threadIdx.x < x.nranks) {
        __threadfence_system();                              // cumulative over every CTA's stores (tickets + block barrier): ordered before the flag
        asm volatile("st.relaxed.sys.global.u32 [%0], %1;\n" ::"l"(flags_of(x, threadIdx.x) + which + x.rank), "r"(x.epoch) : "memory");
    }
}
// steps 2 / 4, second half: the whole CTA waits until all R ranks have published flag `which`
__device__ __forceinline__ void gather_wait(const SgpXchg& x, int which) {
    if (threadIdx.x < x.nranks) wait_epoch(flags_of(x, x.rank) + which + threadIdx.x, x.epoch);
    __syncthreads();
}

// step 3: this rank's share of the n packed elements (pairs of doubles; the buffers are 16-byte aligned and padded to an even length)
__device__ __forceinline__ void reduce_scatter(const SgpXchg& x, long long n, int cta, int ncta) {
    const int R = x.nranks;
    const long long pairs = (n + 1) / 2, p0 = pairs * x.rank / R, p1 = pairs * (x.rank + 1) / R;
    const long long stride = (long long)ncta * blockDim.x;
    for (long long e = p0 + (long long)cta * blockDim.x + threadIdx.x; e < p1; e += stride) {
        double2 v[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) if (q < R) v[q] = __ldcv(reinterpret_cast<const double2*>(contrib_of(x, q)) + e);
        double2 s = v[0];
#pragma unroll
        for (int q = 1; q < 8; ++q) if (q < R) { s.x += v[q].x; s.y += v[q].y; }
#pragma unroll
        for (int q = 0; q < 8; ++q) if (q < R) reinterpret_cast<double2*>(result_of(x, q))[e] = s;
    }
}

// step 5, vector form: dst[e] = result[e]
__device__ __forceinline__ void expand_vec(const SgpXchg& x, double* __restrict__ dst, long long n, int cta, int ncta) {
    const double* res = result_of(x, x.rank);
    for (long long e = (long long)cta * blockDim.x + threadIdx.x; e < n; e += (long long)ncta * blockDim.x) dst[e] = __ldcg(res + e);
}
// step 5, statistics form: packed result -> full symmetric Psi2 (M x M) | tail (ntail doubles) at `stats`
__device__ __forceinline__ void expand_stats(const SgpXchg& x, double* __restrict__ stats, int M, int ntail, int cta, int ncta) {
    const double* res = result_of(x, x.rank);
    const long long tri = (long long)M * (M + 1) / 2, total = tri + ntail;
    const double b = 2.0 * M + 1.0;
    for (long long p = (long long)cta * blockDim.x + threadIdx.x; p < total; p += (long long)ncta * blockDim.x) {
        const double s = __ldcg(res + p);                  // (L2 is the coherence point of the peers' stores: bypass L1)
        if (p >= tri) { stats[(size_t)M * M + (p - tri)] = s; continue; }
        // column j of packed index p: tri_col(j) <= p < tri_col(j + 1)
        long long j = (long long)((b - sqrt(b * b - 8.0 * (double)p)) * 0.5);
        j = j < 0 ? 0 : (j > M - 1 ? M - 1 : j);
        while (tri_col(j, M) > p) --j;
        while (tri_col(j + 1, M) <= p) ++j;
        const long long i = j + (p - tri_col(j, M));
        stats[(size_t)i + (size_t)j * M] = s;
        stats[(size_t)j + (size_t)i * M] = s;
    }
}

// steps 2 - 5 for the statistics (the contribution is in place)
// (t, optional: five clock stamps after publish A / wait A / reduce-scatter / publish B + wait B / expand, for tools/xchg_bench.py)
__device__ __forceinline__ void allreduce_stats(const SgpXchg& x, double* __restrict__ stats, int M, int ntail, int cta, int ncta, long long* t = nullptr) {
    const long long total = (long long)M * (M + 1) / 2 + ntail;
    publish(x, ncta, kFlagA);
    if (t) t[0] = clock64();
    gather_wait(x, kFlagA);
    if (t) t[1] = clock64();
    reduce_scatter(x, total, cta, ncta);
    if (t) t[2] = clock64();
    publish(x, ncta, kFlagB);
    gather_wait(x, kFlagB);
    if (t) t[3] = clock64();
    expand_stats(x, stats, M, ntail, cta, ncta);
    if (t) t[4] = clock64();
}

}  // namespace sgp_xchg
