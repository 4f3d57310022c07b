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
    a.GetErrorString = (decltype(a.GetErrorString))dlsym(a.handle, "ncclGetErrorString");
    a.ok = a.GetUniqueId && a.CommInitRank && a.CommDestroy && a.AllReduce;
    return a;
}
}  // namespace

struct SgpComm {
    ncclComm_t comm = nullptr;
    int nranks = 1, rank = 0;
};

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
    return SGP_OK;
}

int sgp_comm_allreduce(sgp_ctx* ctx, double* buf, size_t count) {
    if (!ctx->comm) return SGP_OK;
    NcclApi& a = api();
    ncclResult_t r = a.AllReduce(buf, buf, count, ncclFloat64, ncclSum, ctx->comm->comm, ctx->stream);
    if (r != 0) {
        ctx->err = std::string("ncclAllReduce: ") + (a.GetErrorString ? a.GetErrorString(r) : "error");
        return SGP_ERR_COMM;
    }
    return SGP_OK;
}

void sgp_comm_destroy(sgp_ctx* ctx) {
    if (ctx && ctx->comm) {
        NcclApi& a = api();
        if (a.ok && ctx->comm->comm) a.CommDestroy(ctx->comm->comm);
        delete ctx->comm;
        ctx->comm = nullptr;
    }
}
