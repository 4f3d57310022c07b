// The peer-memory sum over the ranks (xchg.cuh) as a kernel of its own: for the statistics of the sweeps that do not carry the exchange in
// their own tail (first fused kernel for M <= 384, uncertain-input sweeps: GPnode/MultiSGPnode.jl:290-328 summed over the nodes of ALL
// ranks) and for small vectors (the rank-local part of the theta gradient).  Falls back to one ncclAllReduce when the regions are not mapped.
#include "xchg.cuh"
#include <algorithm>

namespace {

// mode 0: dst[e] = sum over ranks of src[e], e < count.
// mode 1: src = dst = statistics buffer [Psi2 (M x M, full symmetric) | tail (ntail)]: contributed as the packed lower triangle + tail
//         (in place: every CTA has copied its part out before any CTA passes the first gather_wait, which needs this rank's own flag too).
__global__ void __launch_bounds__(256) xchg_kernel(const SgpXchg x, const double* __restrict__ src, double* __restrict__ dst, long long count, int M,
                                                   int ntail, int mode) {
    const int cta = blockIdx.x, ncta = gridDim.x, tid = threadIdx.x, nt = blockDim.x;
    if (mode == 0) {
        for (long long e = (long long)cta * nt + tid; e < count; e += (long long)ncta * nt) sgp_xchg::put1(x, e, src[e]);
        if (cta == 0 && tid == 0 && (count & 1)) sgp_xchg::put1(x, count, 0.0);       // (the reduce-scatter works on pairs)
        sgp_xchg::publish(x, ncta, sgp_xchg::kFlagA);
        sgp_xchg::gather_wait(x, sgp_xchg::kFlagA);
        sgp_xchg::reduce_scatter(x, count, cta, ncta);
        sgp_xchg::publish(x, ncta, sgp_xchg::kFlagB);
        sgp_xchg::gather_wait(x, sgp_xchg::kFlagB);
        sgp_xchg::expand_vec(x, dst, count, cta, ncta);
    } else {
        for (int j = cta; j < M; j += ncta) {
            const long long off = sgp_xchg::tri_col(j, M) - j;
            for (int i = j + tid; i < M; i += nt) sgp_xchg::put1(x, off + i, src[(size_t)i + (size_t)j * M]);
        }
        if (cta == ncta - 1) {
            const long long tri = (long long)M * (M + 1) / 2;
            for (int e = tid; e < ntail; e += nt) sgp_xchg::put1(x, tri + e, src[(size_t)M * M + e]);
            if (tid == 0 && ((tri + ntail) & 1)) sgp_xchg::put1(x, tri + ntail, 0.0);
        }
        sgp_xchg::allreduce_stats(x, dst, M, ntail, cta, ncta);
    }
}

// Flag barrier over the ranks, nothing else (its own flag words and epoch): aligns the ranks in front of a timed region.
__global__ void xchg_barrier_kernel(const SgpXchg x) {
    constexpr int kFlagC = 48;
    const int q = threadIdx.x;
    if (q < x.nranks) {
        __threadfence_system();
        asm volatile("st.relaxed.sys.global.u32 [%0], %1;\n" ::"l"(sgp_xchg::flags_of(x, q) + kFlagC + x.rank), "r"(x.epoch) : "memory");
        sgp_xchg::wait_epoch(sgp_xchg::flags_of(x, x.rank) + kFlagC + q, x.epoch);
    }
}

int launch(sgp_ctx* ctx, const SgpXchg& x, const double* src, double* dst, long long count, int M, int ntail, int mode) {
    int grid = mode == 1 ? std::min(ctx->num_sms, std::max(1, M)) : (int)std::min<long long>(ctx->num_sms, std::max<long long>(1, (count + 1023) / 1024));
    // cooperative launch: the CTAs wait for each other (ticket counter + flags), so they must all be resident
    void* args[] = {(void*)&x, (void*)&src, (void*)&dst, (void*)&count, (void*)&M, (void*)&ntail, (void*)&mode};
    SGP_CUDA(ctx, cudaLaunchCooperativeKernel((const void*)xchg_kernel, dim3(grid), dim3(256), args, 0, ctx->stream));
    SGP_CUDA(ctx, cudaGetLastError());
    return SGP_OK;
}

}  // namespace

int sgp_comm_allreduce_nccl(sgp_ctx* ctx, double* buf, size_t count);

// buf[0 .. count) summed over the ranks, in place
int sgp_comm_allreduce(sgp_ctx* ctx, double* buf, size_t count) {
    if (!ctx->comm) return SGP_OK;
    SgpXchg x;
    if (sgp_comm_xchg(ctx, count, &x)) return launch(ctx, x, buf, buf, (long long)count, 0, 0, 0);
    return sgp_comm_allreduce_nccl(ctx, buf, count);
}

// the resident statistics [Psi2 (M x M) | Psi1 (M x D_out) | 4 scalars] summed over the ranks, in place (bitwise identical on all ranks)
int sgp_comm_allreduce_stats(sgp_ctx* ctx, int M, int D_out) {
    if (!ctx->comm) return SGP_OK;
    SGP_RANGE("sgp_exchange");
    const int ntail = M * D_out + 4;
    const size_t full = (size_t)M * M + (size_t)ntail;
    SgpXchg x;
    if (sgp_comm_xchg(ctx, (size_t)M * (M + 1) / 2 + (size_t)ntail, &x)) return launch(ctx, x, ctx->stats_dev, ctx->stats_dev, 0, M, ntail, 1);
    return sgp_comm_allreduce_nccl(ctx, ctx->stats_dev, full);
}

// Enqueues a barrier over the ranks on the ctx stream (no-op without the peer-memory exchange)
bool sgp_comm_xchg_barrier(sgp_ctx* ctx, SgpXchg* x);
int sgp_comm_barrier(sgp_ctx* ctx) {
    SgpXchg x;
    if (!ctx->comm || !sgp_comm_xchg_barrier(ctx, &x)) return SGP_OK;
    xchg_barrier_kernel<<<1, 32, 0, ctx->stream>>>(x);
    SGP_CUDA(ctx, cudaGetLastError());
    return SGP_OK;
}
