// ps_comm.cu -- the one exchange step of the path: splitting ONE region's events across the GPUs of a box
// (SURVEY.md 8e level 2; poreseq variant on deep coverage, poreseq/Variant.py:71-76).  score[m] = -1e-6 + sum over
// events of delta(m, e) (cpp/MakeMutations.cpp:19-22, 38-52): every rank scores all mutations against ITS block of
// events, the per-mutation sums are combined over NCCL (NVLink / NVSwitch) on the context's stream, behind the kernels,
// with no host round trip and no torch in the data path.
//
// Two combine modes (ps_comm_init's `ordered`):
//   ordered   the reference adds the events of a region in event order, in FP64.  Rank r receives the running sums of
//             ranks < r (ncclRecv), continues them over its own events in order (k_reduce with a start array) and sends
//             them on (ncclSend); the last rank broadcasts the totals.  Bit-identical to the single-GPU path and to the
//             reference; costs nranks - 1 small messages in sequence (n_mutations doubles each).
//   allreduce one ncclAllReduce(sum, float64, n_mutations) of the ranks' partial sums (what BASELINE.json's north_star
//             names): scores agree to ~1e-16 relative, accept / reject of clearly signed scores identical.
//
// NCCL is loaded with dlopen at ps_comm_init (libnccl.so.2: the copy a host program such as torch already has mapped,
// else the system's), so the library itself has no link-time dependency on it and single-GPU users never touch it.
#include <dlfcn.h>
#include <nccl.h>

#include <cstdio>
#include <cstdlib>
#include <mutex>

#include "ps_internal.h"

namespace
{
struct NcclApi
{
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    std::string error;
};

NcclApi* nccl_api()
{
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        const char* names[] = {getenv("PORESEQ_B200_NCCL"), "libnccl.so.2", "libnccl.so"};
        for (const char* n : names)
        {
            if (!n || !*n) continue;
            api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (api.handle) break;
        }
        if (!api.handle) { api.error = "libnccl.so.2 could not be loaded (set PORESEQ_B200_NCCL to its path)"; return; }
#define PS_NCCL_SYM(field, name)                                                           \
    api.field = reinterpret_cast<decltype(api.field)>(dlsym(api.handle, name));            \
    if (!api.field && api.error.empty()) api.error = std::string("libnccl has no ") + name;
        PS_NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
        PS_NCCL_SYM(CommInitRank, "ncclCommInitRank")
        PS_NCCL_SYM(CommDestroy, "ncclCommDestroy")
        PS_NCCL_SYM(AllReduce, "ncclAllReduce")
        PS_NCCL_SYM(Broadcast, "ncclBroadcast")
        PS_NCCL_SYM(Send, "ncclSend")
        PS_NCCL_SYM(Recv, "ncclRecv")
        PS_NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef PS_NCCL_SYM
    });
    return &api;
}
} // namespace

#define NC(call)                                                                                          \
    do {                                                                                                  \
        ncclResult_t r__ = (call);                                                                        \
        if (r__ != ncclSuccess)                                                                           \
        {                                                                                                 \
            ps_set_error(ctx, "NCCL error %s at %s:%d (%s)", nccl_api()->GetErrorString(r__), __FILE__, __LINE__, #call); \
            return PS_E_CUDA;                                                                             \
        }                                                                                                 \
    } while (0)

// ---- used by Job::run (ps_host.cu) ----------------------------------------------------------------------
int psi_comm_allreduce_sum(ps_ctx* ctx, double* buf, size_t count)
{
    NC(nccl_api()->AllReduce(buf, buf, count, ncclFloat64, ncclSum, (ncclComm_t)ctx->comm, ctx->stream));
    return PS_OK;
}
int psi_comm_recv_prev(ps_ctx* ctx, double* buf, size_t count)
{
    NC(nccl_api()->Recv(buf, count, ncclFloat64, ctx->comm_rank - 1, (ncclComm_t)ctx->comm, ctx->stream));
    return PS_OK;
}
int psi_comm_send_next(ps_ctx* ctx, const double* buf, size_t count)
{
    NC(nccl_api()->Send(buf, count, ncclFloat64, ctx->comm_rank + 1, (ncclComm_t)ctx->comm, ctx->stream));
    return PS_OK;
}
int psi_comm_bcast_last(ps_ctx* ctx, double* buf, size_t count)
{
    NC(nccl_api()->Broadcast(buf, buf, count, ncclFloat64, ctx->comm_ranks - 1, (ncclComm_t)ctx->comm, ctx->stream));
    return PS_OK;
}

extern "C" {

int ps_comm_unique_id(void* id, int bytes)
{
    NcclApi* a = nccl_api();
    if (!id || bytes < (int)sizeof(ncclUniqueId)) return PS_BAD_ARGS(nullptr, "ps_comm_unique_id");
    if (!a->error.empty()) { ps_set_error(nullptr, "%s", a->error.c_str()); return PS_E_CUDA; }
    ncclUniqueId u;
    if (a->GetUniqueId(&u) != ncclSuccess) { ps_set_error(nullptr, "ncclGetUniqueId failed"); return PS_E_CUDA; }
    memcpy(id, &u, sizeof u);
    return PS_OK;
}

int ps_comm_init(ps_ctx* ctx, const void* id, int bytes, int rank, int n_ranks, int ordered)
{
    if (!ctx || !id || bytes < (int)sizeof(ncclUniqueId) || n_ranks < 1 || rank < 0 || rank >= n_ranks)
        return PS_BAD_ARGS(ctx, "ps_comm_init");
    NcclApi* a = nccl_api();
    if (!a->error.empty()) { ps_set_error(ctx, "%s", a->error.c_str()); return PS_E_CUDA; }
    if (ctx->comm) { ps_set_error(ctx, "ps_comm_init: this context already belongs to a communicator"); return PS_E_ARG; }
    int rc = ctx->init();
    if (rc) return rc;
    if (cudaSetDevice(ctx->device) != cudaSuccess) { ps_set_error(ctx, "ps_comm_init: cudaSetDevice failed"); return PS_E_CUDA; }
    ncclUniqueId u;
    memcpy(&u, id, sizeof u);
    ncclComm_t comm = nullptr;
    NC(a->CommInitRank(&comm, n_ranks, u, rank));
    ctx->comm = comm; ctx->comm_rank = rank; ctx->comm_ranks = n_ranks; ctx->comm_ordered = ordered != 0;
    return PS_OK;
}

int ps_comm_destroy(ps_ctx* ctx)
{
    if (!ctx) return PS_BAD_ARGS(ctx, "ps_comm_destroy");
    if (ctx->comm)
    {
        if (ctx->stream) cudaStreamSynchronize(ctx->stream);
        nccl_api()->CommDestroy((ncclComm_t)ctx->comm);
        ctx->comm = nullptr; ctx->comm_rank = 0; ctx->comm_ranks = 1;
    }
    return PS_OK;
}

int ps_comm_rank(ps_ctx* ctx, int* rank, int* n_ranks)
{
    if (!ctx) return PS_BAD_ARGS(ctx, "ps_comm_rank");
    if (rank) *rank = ctx->comm_rank;
    if (n_ranks) *n_ranks = ctx->comm_ranks;
    return PS_OK;
}

} // extern "C"
