// Multi-GPU exchange of the sweep: N is sharded over ranks (one sgp_ctx per rank/GPU); the only data-path collective
// is ONE all-reduce (sum, f64) of the packed statistics [Psi2 | Psi1 | Psi0 | sum_y2 | sum_w | n] per sweep, over
// NVLink 5 / NVSwitch via NCCL.  The reference has no distributed code at all; this is new.
// NCCL is resolved at run time with dlopen so that the library also loads in a process that already carries torch's
// bundled libnccl (same soname), and on hosts without NCCL (single-GPU use).
#include "sgp_internal.cuh"
#include <dlfcn.h>
#include <cstring>

namespace {
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclFloat64 = 8, ncclSum = 0 };

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};

NcclApi& api() {
    static NcclApi a;
    if (a.handle || a.ok) return a;
    const char* names[] = {"libnccl.so.2", "libnccl.so", nullptr};
    for (int i = 0; names[i] && !a.handle; ++i) a.handle = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL);
    if (!a.handle) return a;
    a.GetUniqueId = (decltype(a.GetUniqueId))dlsym(a.handle, "ncclGetUniqueId");
    a.CommInitRank = (decltype(a.CommInitRank))dlsym(a.handle, "ncclCommInitRank");
    a.CommDestroy = (decltype(a.CommDestroy))dlsym(a.handle, "ncclCommDestroy");
    a.AllReduce = (decltype(a.AllReduce))dlsym(a.handle, "ncclAllReduce");
    a.AllGather = (decltype(a.AllGather))dlsym(a.handle, "ncclAllGather");
    a.GetErrorString = (decltype(a.GetErrorString))dlsym(a.handle, "ncclGetErrorString");
    a.ok = a.GetUniqueId && a.CommInitRank && a.CommDestroy && a.AllReduce;
    return a;
}
}  // namespace

struct SgpComm {
    ncclComm_t comm = nullptr;
    int nranks = 1, rank = 0;
    // fused peer-memory exchange (see SgpXchg): own region + the peers' regions mapped through CUDA IPC
    bool p2p = false;
    char* region = nullptr;
    size_t cap_doubles = 0;
    char* peers[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    unsigned epoch = 0;
    unsigned epoch_bar = 0;        // epoch of the stand-alone rank barrier (sgp_comm_barrier)
};

namespace {
constexpr size_t kFlagBytes = 256;      // flags A [8] at 0, flags B [8] at 64

// Collective over the communicator: allocate the exchange region, trade IPC handles through ncclAllGather, map the peers.
// Any failure (ranks on different nodes, no peer access, ...) leaves p2p off on EVERY rank: the success flags are summed.
void setup_p2p(sgp_ctx* ctx, SgpComm* c) {
    NcclApi& a = api();
    if (c->nranks < 2 || c->nranks > 8 || !a.AllGather) return;
    if (const char* e = std::getenv("SGP_COMM_P2P")) if (e[0] == '0') return;
    size_t maxM = 2048;
    if (const char* e = std::getenv("SGP_COMM_P2P_MAXM")) { long v = std::atol(e); if (v >= 1) maxM = (size_t)v; }
    // two buffers (contribution, result) of the packed statistics (lower triangle of Psi2 | Psi1 for up to 16 outputs | scalars), 16-byte aligned
    const size_t cap = (maxM * (maxM + 1) / 2 + 16 * maxM + 64 + 1) & ~(size_t)1;
    const size_t bytes = kFlagBytes + 2 * cap * sizeof(double);
    bool ok = cudaMalloc((void**)&c->region, bytes) == cudaSuccess;
    if (ok) ok = cudaMemset(c->region, 0, bytes) == cudaSuccess;
    cudaIpcMemHandle_t mine;
    memset(&mine, 0, sizeof mine);
    if (ok) ok = cudaIpcGetMemHandle(&mine, c->region) == cudaSuccess;
    // handles (64 bytes each) + one success word + the device's PCI identity per rank travel through NCCL
    const size_t rec = sizeof(cudaIpcMemHandle_t) + 16;
    long long pci = -1;
    {
        cudaDeviceProp prop;
        if (cudaGetDeviceProperties(&prop, ctx->dev) == cudaSuccess)
            pci = ((long long)prop.pciDomainID << 32) | ((long long)prop.pciBusID << 16) | (long long)prop.pciDeviceID;
    }
    char* dev = nullptr;
    std::vector<char> host(rec * c->nranks, 0);
    bool ok_all = cudaMalloc((void**)&dev, rec * (c->nranks + 1)) == cudaSuccess;
    if (ok_all) {
        char mine_rec[sizeof(cudaIpcMemHandle_t) + 16];
        memcpy(mine_rec, &mine, sizeof mine);
        long long flag = ok ? 1 : 0;
        memcpy(mine_rec + sizeof mine, &flag, 8);
        memcpy(mine_rec + sizeof mine + 8, &pci, 8);
        ok_all = cudaMemcpyAsync(dev + rec * c->nranks, mine_rec, rec, cudaMemcpyHostToDevice, ctx->stream) == cudaSuccess;
        if (a.AllGather(dev + rec * c->nranks, dev, rec, /*ncclChar*/ 0, c->comm, ctx->stream) != 0) ok_all = false;
        if (cudaMemcpyAsync(host.data(), dev, rec * c->nranks, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess) ok_all = false;
        if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) ok_all = false;
    }
    if (dev) cudaFree(dev);
    int good = 0;
    if (ok_all)
        for (int q = 0; q < c->nranks; ++q) { long long f = 0; memcpy(&f, host.data() + rec * q + sizeof(cudaIpcMemHandle_t), 8); good += f == 1; }
    bool mapped = ok_all && good == c->nranks;
    // two ranks on ONE device cannot run their sweep kernels side by side (time-sliced contexts): the in-kernel barriers would never meet
    if (mapped)
        for (int a_ = 0; a_ < c->nranks && mapped; ++a_)
            for (int b_ = a_ + 1; b_ < c->nranks; ++b_) {
                long long pa = 0, pb = 0;
                memcpy(&pa, host.data() + rec * a_ + sizeof(cudaIpcMemHandle_t) + 8, 8);
                memcpy(&pb, host.data() + rec * b_ + sizeof(cudaIpcMemHandle_t) + 8, 8);
                if (pa == pb || pa < 0 || pb < 0) { mapped = false; break; }
            }
    if (mapped) {
        for (int q = 0; q < c->nranks && mapped; ++q) {
            if (q == c->rank) { c->peers[q] = c->region; continue; }
            cudaIpcMemHandle_t h;
            memcpy(&h, host.data() + rec * q, sizeof h);
            void* ptr = nullptr;
            if (cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); mapped = false; }
            c->peers[q] = (char*)ptr;
        }
    }
    // second agreement round: every rank must have mapped every peer
    int* dflag = nullptr;
    int agreed = 0;
    if (cudaMalloc((void**)&dflag, sizeof(int) * (c->nranks + 1)) == cudaSuccess) {
        int v = mapped ? 1 : 0;
        cudaMemcpyAsync(dflag + c->nranks, &v, sizeof v, cudaMemcpyHostToDevice, ctx->stream);
        std::vector<int> all(c->nranks, 0);
        if (a.AllGather(dflag + c->nranks, dflag, sizeof(int), 0, c->comm, ctx->stream) == 0 &&
            cudaMemcpyAsync(all.data(), dflag, sizeof(int) * c->nranks, cudaMemcpyDeviceToHost, ctx->stream) == cudaSuccess &&
            cudaStreamSynchronize(ctx->stream) == cudaSuccess)
            for (int q = 0; q < c->nranks; ++q) agreed += all[q] == 1;
        cudaFree(dflag);
    }
    if (agreed == c->nranks) { c->p2p = true; c->cap_doubles = cap; return; }
    for (int q = 0; q < c->nranks; ++q) if (q != c->rank && c->peers[q]) { cudaIpcCloseMemHandle(c->peers[q]); c->peers[q] = nullptr; }
    if (c->region) { cudaFree(c->region); c->region = nullptr; }
}
}  // namespace

extern "C" int sgp_comm_unique_id(char id[128]) {
    if (!id) return SGP_ERR_ARG;
    NcclApi& a = api();
    if (!a.ok) return SGP_ERR_COMM;
    ncclUniqueId u;
    if (a.GetUniqueId(&u) != 0) return SGP_ERR_COMM;
    memcpy(id, u.internal, 128);
    return SGP_OK;
}

extern "C" int sgp_comm_init(sgp_ctx* ctx, int nranks, int rank, const char id[128]) {
    if (!ctx) return SGP_ERR_ARG;
    if (nranks < 1 || rank < 0 || rank >= nranks || !id) SGP_FAIL(ctx, SGP_ERR_ARG, "comm_init: bad rank / nranks / id");
    NcclApi& a = api();
    if (!a.ok) SGP_FAIL(ctx, SGP_ERR_COMM, "comm_init: libnccl.so.2 not found");
    SGP_CUDA(ctx, cudaSetDevice(ctx->dev));
    sgp_comm_destroy(ctx);
    SgpComm* c = new SgpComm();
    ncclUniqueId u;
    memcpy(u.internal, id, 128);
    ncclResult_t r = a.CommInitRank(&c->comm, nranks, u, rank);
    if (r != 0) {
        ctx->err = std::string("ncclCommInitRank: ") + (a.GetErrorString ? a.GetErrorString(r) : "error");
        delete c;
        return SGP_ERR_COMM;
    }
    c->nranks = nranks; c->rank = rank;
    ctx->comm = c;
    setup_p2p(ctx, c);
    return SGP_OK;
}

int sgp_comm_allreduce_nccl(sgp_ctx* ctx, double* buf, size_t count) {
    if (!ctx->comm) return SGP_OK;
    NcclApi& a = api();
    ncclResult_t r = a.AllReduce(buf, buf, count, ncclFloat64, ncclSum, ctx->comm->comm, ctx->stream);
    if (r != 0) {
        ctx->err = std::string("ncclAllReduce: ") + (a.GetErrorString ? a.GetErrorString(r) : "error");
        return SGP_ERR_COMM;
    }
    return SGP_OK;
}

bool sgp_comm_xchg(sgp_ctx* ctx, size_t need_doubles, SgpXchg* x) {
    SgpComm* c = ctx->comm;
    if (!c || !c->p2p || need_doubles > c->cap_doubles) return false;
    x->nranks = c->nranks; x->rank = c->rank; x->epoch = ++c->epoch;
    ctx->packed_src = nullptr;      // the result buffer is about to be rewritten: a packed copy the host was pointed at (sweep.cu) is no longer there
    for (int q = 0; q < 8; ++q) x->peers[q] = c->peers[q];
    x->slot0_off = kFlagBytes; x->slot_bytes = c->cap_doubles * sizeof(double);
    return true;
}

bool sgp_comm_xchg_barrier(sgp_ctx* ctx, SgpXchg* x) {
    SgpComm* c = ctx->comm;
    if (!c || !c->p2p) return false;
    x->nranks = c->nranks; x->rank = c->rank; x->epoch = ++c->epoch_bar;
    for (int q = 0; q < 8; ++q) x->peers[q] = c->peers[q];
    return true;
}

void sgp_comm_destroy(sgp_ctx* ctx) {
    if (ctx && ctx->comm) {
        NcclApi& a = api();
        SgpComm* c = ctx->comm;
        if (ctx->stream) cudaStreamSynchronize(ctx->stream);
        for (int q = 0; q < 8; ++q) if (q != c->rank && c->peers[q]) cudaIpcCloseMemHandle(c->peers[q]);
        if (c->region) cudaFree(c->region);
        if (a.ok && ctx->comm->comm) a.CommDestroy(ctx->comm->comm);
        delete ctx->comm;
        ctx->comm = nullptr;
    }
}
