// Sum of the per-rank statistics over NVLink / NVSwitch peer memory, as device code that runs in the TAIL of the kernel that produced them
// (sweep4_kernel.cuh) or as a small kernel of its own (xchg.cu) -- no host involvement, no NCCL launch.  The reference has no
// distributed code; this is the one collective of the N-sharded sweep (section 8e).
//
// Every rank owns a region [flags (256 B) | xin (cap doubles) | xout (cap doubles)] that all peers map through CUDA IPC (comm.cu).
// ONE-SHOT PULL for R <= 8 ranks:
//   1. the rank writes its contribution to its OWN xin -- for the sweep statistics in PACKED form: the lower triangle of Psi2 column by
//      column (M (M + 1) / 2 doubles), then Psi1 and the four scalars.  Before that it waits until every peer has finished reading the
//      previous contribution (flag B of the previous epoch; nobody waits for that at the time it is signalled).
//   2. the LAST CTA to finish (ticket counter, no grid barrier) publishes flag A = epoch into every rank's flag array (st.release.sys).
//   3. one warp per CTA polls the local flag array until all R ranks have published, then every CTA pulls its share of ALL R contributions
//      (ld.cv over NVLink), adds them in rank order -- the same order on every rank: bitwise identical sums -- and writes the full
//      symmetric Psi2 / Psi1 / scalars into the local statistics buffer.
//   4. the last CTA to finish signals flag B = epoch to every peer.  The kernel ends without waiting for anybody.
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
__device__ __forceinline__ double* xin_of(const SgpXchg& x, int q) { return reinterpret_cast<double*>(x.peers[q] + x.xin_off); }

// packed offset of column j of an M x M lower triangle stored column by column
__device__ __forceinline__ long long tri_col(long long j, long long M) { return j * M - j * (j - 1) / 2; }

// step 1 (entry): the whole CTA waits until every peer has read this rank's previous contribution
__device__ __forceinline__ void wait_free(const SgpXchg& x) {
    if (threadIdx.x < x.nranks) wait_epoch(flags_of(x, x.rank) + kFlagB + threadIdx.x, x.epoch - 1u);
    __syncthreads();
}
// step 2: this CTA's part of xin is written; the last of the `ncta` CTAs publishes the contribution
__device__ __forceinline__ void publish(const SgpXchg& x, int ncta) {
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned* cnt = flags_of(x, x.rank) + kCntA;
        if (atomicAdd(cnt, 1u) == (unsigned)ncta - 1u) {
            *cnt = 0u;
            __threadfence_system();                          // cumulative: every CTA's xin stores are ordered before the flags below
            for (int q = 0; q < x.nranks; ++q) st_release_sys(flags_of(x, q) + kFlagA + x.rank, x.epoch);
        }
    }
}
// step 3a: wait for all contributions
__device__ __forceinline__ void gather_wait(const SgpXchg& x) {
    if (threadIdx.x < x.nranks) wait_epoch(flags_of(x, x.rank) + kFlagA + threadIdx.x, x.epoch);
    __syncthreads();
}
// step 4
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

// sum over the ranks (rank order) of element e of the contributions
__device__ __forceinline__ double pull1(const SgpXchg& x, long long e) {
    double v[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) if (q < x.nranks) v[q] = __ldcv(xin_of(x, q) + e);
    double s = v[0];
#pragma unroll
    for (int q = 1; q < 8; ++q) if (q < x.nranks) s += v[q];
    return s;
}

// step 3b, statistics form: packed contributions -> full symmetric Psi2 (M x M) | tail (ntail doubles) at `stats`.
// CTA c of ncta takes the columns c, c + ncta, ...; two elements per thread are in flight.
__device__ __forceinline__ void pull_stats(const SgpXchg& x, double* __restrict__ stats, int M, int ntail, int cta, int ncta) {
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int j = cta; j < M; j += ncta) {
        const long long off = tri_col(j, M) - j;                    // packed index of (i, j) = off + i
        for (int i = j + tid; i < M; i += 2 * nt) {
            const int i2 = i + nt;
            double v[8], u[8];
#pragma unroll
            for (int q = 0; q < 8; ++q)
                if (q < x.nranks) {
                    v[q] = __ldcv(xin_of(x, q) + off + i);
                    if (i2 < M) u[q] = __ldcv(xin_of(x, q) + off + i2);
                }
            double s = v[0], s2 = i2 < M ? u[0] : 0.0;
#pragma unroll
            for (int q = 1; q < 8; ++q) if (q < x.nranks) { s += v[q]; if (i2 < M) s2 += u[q]; }
            stats[(size_t)i + (size_t)j * M] = s;
            stats[(size_t)j + (size_t)i * M] = s;
            if (i2 < M) { stats[(size_t)i2 + (size_t)j * M] = s2; stats[(size_t)j + (size_t)i2 * M] = s2; }
        }
    }
    if (cta == ncta - 1) {
        const long long tri = (long long)M * (M + 1) / 2;
        for (int e = tid; e < ntail; e += nt) stats[(size_t)M * M + e] = pull1(x, tri + e);
    }
}

}  // namespace sgp_xchg
