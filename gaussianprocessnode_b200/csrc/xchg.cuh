// Sum of the per-rank statistics over NVLink / NVSwitch peer memory, as device code that runs in the TAIL of the kernel that produced them
// (sweep4_kernel.cuh) or as a small kernel of its own (xchg.cu) -- no host involvement, no NCCL launch.  The reference has no
// distributed code; this is the one collective of the N-sharded sweep (section 8e).
//
// Every rank owns a region [flags (256 B) | slot 0 | ... | slot R-1] that all peers map through CUDA IPC (comm.cu); slot q of rank r
// receives rank q's contribution.  ONE-SHOT PUSH for R <= 8 ranks:
//   1. before it writes, the rank waits until every peer has finished reading the previous contribution (flag B of the previous epoch;
//      nobody waits for that at the time it is signalled).
//   2. the rank PUSHES its contribution into slot `rank` of EVERY rank (plain stores over NVLink, fire and forget: the wire time hides
//      under the phase that produces the values) -- for the sweep statistics in PACKED form: the lower triangle of Psi2 column by column
//      (M (M + 1) / 2 doubles), then Psi1 and the four scalars.
//   3. the LAST CTA to finish (ticket counter, no grid barrier) publishes flag A = epoch into every rank's flag array (st.release.sys after a
//      system fence).
//   4. one warp per CTA polls the LOCAL flag array until all R ranks have published; then every CTA adds its share of the R local slots in
//      rank order -- the same order on every rank: bitwise identical sums -- and writes the full symmetric Psi2 / Psi1 / scalars into the
//      local statistics buffer.  No remote read anywhere.
//   5. the last CTA to finish signals flag B = epoch to every peer.  The kernel ends without waiting for anybody.
// Flags are monotonic epochs, never reset; the two ticket counters are reset by the CTA that completes them.
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
    if ((int)(ld_acquire_sys(flag) - epoch) >= 0) return;
    const long long t0 = clock64();
    while ((int)(ld_acquire_sys(flag) - epoch) < 0) {
        __nanosleep(40);
        if (clock64() - t0 > 40000000000ll) __trap();
    }
}
__device__ __forceinline__ unsigned* flags_of(const SgpXchg& x, int q) { return reinterpret_cast<unsigned*>(x.peers[q]); }
// slot `src` of rank `dst`: where rank src's contribution lands on rank dst
__device__ __forceinline__ double* slot_of(const SgpXchg& x, int dst, int src) { return reinterpret_cast<double*>(x.peers[dst] + x.slot0_off + (size_t)src * x.slot_bytes); }
// element e of this rank's contribution, pushed to every rank
__device__ __forceinline__ void push1(const SgpXchg& x, long long e, double v) {
#pragma unroll
    for (int q = 0; q < 8; ++q) if (q < x.nranks) slot_of(x, q, x.rank)[e] = v;
}

// packed offset of column j of an M x M lower triangle stored column by column
__device__ __forceinline__ long long tri_col(long long j, long long M) { return j * M - j * (j - 1) / 2; }

// step 1 (entry): the whole CTA waits until every peer has read this rank's previous contribution out of its slot
__device__ __forceinline__ void wait_free(const SgpXchg& x) {
    if (threadIdx.x < x.nranks) wait_epoch(flags_of(x, x.rank) + kFlagB + threadIdx.x, x.epoch - 1u);
    __syncthreads();
}
// step 3: this CTA's part of the contribution is pushed; the last of the `ncta` CTAs publishes it
__device__ __forceinline__ void publish(const SgpXchg& x, int ncta) {
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned* cnt = flags_of(x, x.rank) + kCntA;
        if (atomicAdd(cnt, 1u) == (unsigned)ncta - 1u) {
            *cnt = 0u;
            __threadfence_system();                          // cumulative: every CTA's pushed stores are ordered before the flags below
            for (int q = 0; q < x.nranks; ++q) st_release_sys(flags_of(x, q) + kFlagA + x.rank, x.epoch);
        }
    }
}
// step 4a: wait for all contributions
__device__ __forceinline__ void gather_wait(const SgpXchg& x) {
    if (threadIdx.x < x.nranks) wait_epoch(flags_of(x, x.rank) + kFlagA + threadIdx.x, x.epoch);
    __syncthreads();
}
// step 5
__device__ __forceinline__ void done(const SgpXchg& x, int ncta) {
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned* cnt = flags_of(x, x.rank) + kCntB;
        __threadfence();
        if (atomicAdd(cnt, 1u) == (unsigned)ncta - 1u) {
            *cnt = 0u;
            __threadfence_system();
            for (int q = 0; q < x.nranks; ++q) st_release_sys(flags_of(x, q) + kFlagB + x.rank, x.epoch);
        }
    }
}

// sum over the ranks (rank order) of element e of the contributions received in the LOCAL slots (L2 is the coherence point of peer writes:
// the loads bypass L1)
__device__ __forceinline__ double sum1(const SgpXchg& x, long long e) {
    double v[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) if (q < x.nranks) v[q] = __ldcg(slot_of(x, x.rank, q) + e);
    double s = v[0];
#pragma unroll
    for (int q = 1; q < 8; ++q) if (q < x.nranks) s += v[q];
    return s;
}

// step 4b, statistics form: packed contributions -> full symmetric Psi2 (M x M) | tail (ntail doubles) at `stats`.
// The packed range [0, M (M + 1) / 2 + ntail) is dealt over all threads of the grid, U elements per thread with all their loads in flight.
__device__ __forceinline__ void sum_stats(const SgpXchg& x, double* __restrict__ stats, int M, int ntail, int cta, int ncta) {
    constexpr int U = 4;
    const long long tri = (long long)M * (M + 1) / 2, total = tri + ntail;
    const long long nthreads = (long long)ncta * blockDim.x;
    const double b = 2.0 * M + 1.0;
    for (long long p0 = (long long)cta * blockDim.x + threadIdx.x; p0 < total; p0 += U * nthreads) {
        double v[U][8];
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
            for (int q = 0; q < 8; ++q)
                if (q < x.nranks && p0 + u * nthreads < total) v[u][q] = __ldcg(slot_of(x, x.rank, q) + p0 + u * nthreads);
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long p = p0 + u * nthreads;
            if (p >= total) break;
            double s = v[u][0];
#pragma unroll
            for (int q = 1; q < 8; ++q) if (q < x.nranks) s += v[u][q];
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
}

}  // namespace sgp_xchg
