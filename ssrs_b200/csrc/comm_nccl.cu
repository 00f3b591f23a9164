// NCCL transport behind `ssrs_comm` (include/ssrs_b200.h): halo exchange, scalar all-reduce and all-gather for
// the row-sharded potential solve, and the presence-map all-reduce.  One process per GPU.
//
// libnccl is opened at run time (dlopen "libnccl.so.2"): inside a torch process this resolves to the NCCL that
// torch already loaded, so both share one library; the shared object itself has no link-time NCCL dependency
// and still loads on a machine without NCCL (the entry points below then return SSRS_ERR_UNSUPPORTED).
#include "common.cuh"

#include <dlfcn.h>
#include <nccl.h>
#include <stdlib.h>
#include <string.h>

namespace ssrs {
namespace {

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

NcclApi* api() {
    static NcclApi a;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (h) {
            a.handle = h;
#define SSRS_NCCL_SYM(name) *(void**)(&a.name) = dlsym(h, "nccl" #name)
            SSRS_NCCL_SYM(GetUniqueId); SSRS_NCCL_SYM(CommInitRank); SSRS_NCCL_SYM(CommDestroy); SSRS_NCCL_SYM(Send);
            SSRS_NCCL_SYM(Recv); SSRS_NCCL_SYM(GroupStart); SSRS_NCCL_SYM(GroupEnd); SSRS_NCCL_SYM(AllReduce);
            SSRS_NCCL_SYM(Broadcast); SSRS_NCCL_SYM(GetErrorString);
#undef SSRS_NCCL_SYM
            if (!a.GetUniqueId || !a.CommInitRank || !a.CommDestroy || !a.Send || !a.Recv || !a.GroupStart || !a.GroupEnd ||
                !a.AllReduce || !a.Broadcast)
                a.handle = nullptr;
        }
    }
    return a.handle ? &a : nullptr;
}

struct Ctx {
    ncclComm_t comm = nullptr;
    int rank = 0, size = 1;
    double* dscratch = nullptr;     // kScalars doubles on the device for small reductions (inner products; per-part integers of the setup)
};

int nccl_fail(ncclResult_t r, const char* what) {
    NcclApi* a = api();
    set_error("NCCL error %d (%s) in %s", (int)r, (a && a->GetErrorString) ? a->GetErrorString(r) : "?", what);
    return SSRS_ERR_CUDA;
}
#define SSRS_NCCL_TRY(expr) do { ncclResult_t r_ = (expr); if (r_ != ncclSuccess) return nccl_fail(r_, #expr); } while (0)

constexpr int kScalars = 4 * SSRS_MAX_RANKS;      // the distributed setup exchanges up to two integers per part

int cb_exchange(void* vctx, void* base, int64_t su_off, int64_t su_n, int64_t ru_off, int64_t ru_n,
                int64_t sd_off, int64_t sd_n, int64_t rd_off, int64_t rd_n, void* stream) {
    Ctx* c = (Ctx*)vctx;
    NcclApi* a = api();
    cudaStream_t st = (cudaStream_t)stream;
    char* b = (char*)base;
    if (su_n + ru_n + sd_n + rd_n == 0) return 0;
    SSRS_NCCL_TRY(a->GroupStart());
    if (su_n) SSRS_NCCL_TRY(a->Send(b + su_off, (size_t)su_n, ncclChar, c->rank - 1, c->comm, st));
    if (ru_n) SSRS_NCCL_TRY(a->Recv(b + ru_off, (size_t)ru_n, ncclChar, c->rank - 1, c->comm, st));
    if (sd_n) SSRS_NCCL_TRY(a->Send(b + sd_off, (size_t)sd_n, ncclChar, c->rank + 1, c->comm, st));
    if (rd_n) SSRS_NCCL_TRY(a->Recv(b + rd_off, (size_t)rd_n, ncclChar, c->rank + 1, c->comm, st));
    SSRS_NCCL_TRY(a->GroupEnd());
    return 0;
}

int cb_allreduce_sum(void* vctx, double* host, int32_t count, void* stream) {
    Ctx* c = (Ctx*)vctx;
    NcclApi* a = api();
    cudaStream_t st = (cudaStream_t)stream;
    if (count < 1 || count > kScalars) { set_error("ssrs_comm: all-reduce of %d scalars (1..%d supported)", count, kScalars); return SSRS_ERR_INVALID; }
    SSRS_CUDA_TRY(cudaMemcpyAsync(c->dscratch, host, sizeof(double) * count, cudaMemcpyHostToDevice, st));
    SSRS_NCCL_TRY(a->AllReduce(c->dscratch, c->dscratch, (size_t)count, ncclDouble, ncclSum, c->comm, st));
    SSRS_CUDA_TRY(cudaMemcpyAsync(host, c->dscratch, sizeof(double) * count, cudaMemcpyDeviceToHost, st));
    SSRS_CUDA_TRY(cudaStreamSynchronize(st));
    return 0;
}

int cb_allgather(void* vctx, void* base, const int64_t* offs, void* stream) {
    Ctx* c = (Ctx*)vctx;
    NcclApi* a = api();
    cudaStream_t st = (cudaStream_t)stream;
    char* b = (char*)base;
    SSRS_NCCL_TRY(a->GroupStart());
    for (int r = 0; r < c->size; ++r) {
        const int64_t nbytes = offs[r + 1] - offs[r];
        if (nbytes > 0) SSRS_NCCL_TRY(a->Broadcast(b + offs[r], b + offs[r], (size_t)nbytes, ncclChar, r, c->comm, st));
    }
    SSRS_NCCL_TRY(a->GroupEnd());
    return 0;
}

int cb_allreduce_u32(void* vctx, uint32_t* values, int64_t count, void* stream) {
    Ctx* c = (Ctx*)vctx;
    NcclApi* a = api();
    if (count <= 0) return 0;
    SSRS_NCCL_TRY(a->AllReduce(values, values, (size_t)count, ncclUint32, ncclSum, c->comm, (cudaStream_t)stream));
    return 0;
}

}  // namespace
}  // namespace ssrs

using namespace ssrs;

extern "C" int ssrs_nccl_unique_id(void* id128_host) {
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    NcclApi* a = api();
    if (!a) { set_error("ssrs_nccl_unique_id: libnccl.so.2 not found"); return SSRS_ERR_UNSUPPORTED; }
    if (!id128_host) { set_error("ssrs_nccl_unique_id: NULL buffer"); return SSRS_ERR_INVALID; }
    ncclUniqueId id;
    SSRS_NCCL_TRY(a->GetUniqueId(&id));
    memcpy(id128_host, &id, sizeof(id));
    return SSRS_OK;
}

extern "C" int ssrs_comm_create_nccl(const void* id128_host, int rank, int size, ssrs_comm** out) {
    NcclApi* a = api();
    if (!a) { set_error("ssrs_comm_create_nccl: libnccl.so.2 not found"); return SSRS_ERR_UNSUPPORTED; }
    if (!id128_host || !out || size < 1 || rank < 0 || rank >= size) { set_error("ssrs_comm_create_nccl: bad arguments"); return SSRS_ERR_INVALID; }
    ncclUniqueId id;
    memcpy(&id, id128_host, sizeof(id));
    Ctx* c = new Ctx();
    c->rank = rank; c->size = size;
    ncclResult_t r = a->CommInitRank(&c->comm, size, id, rank);
    if (r != ncclSuccess) { delete c; return nccl_fail(r, "ncclCommInitRank"); }
    if (cudaMalloc(&c->dscratch, kScalars * sizeof(double)) != cudaSuccess) { a->CommDestroy(c->comm); delete c; set_error("ssrs_comm_create_nccl: cudaMalloc failed"); return SSRS_ERR_CUDA; }
    ssrs_comm* m = new ssrs_comm();
    m->rank = rank; m->size = size; m->ctx = c;
    m->exchange = cb_exchange; m->allreduce_sum = cb_allreduce_sum; m->allgather = cb_allgather; m->allreduce_u32 = cb_allreduce_u32;
    *out = m;
    return SSRS_OK;
}

extern "C" int ssrs_comm_destroy(ssrs_comm* m) {
    if (!m) return SSRS_OK;
    NcclApi* a = api();
    Ctx* c = (Ctx*)m->ctx;
    if (c) {
        if (c->dscratch) cudaFree(c->dscratch);
        if (a && c->comm) a->CommDestroy(c->comm);
        delete c;
    }
    delete m;
    return SSRS_OK;
}

extern "C" int ssrs_presence_allreduce(uint32_t* presence, int64_t n, const ssrs_comm* comm, void* stream) {
    if (n < 0 || (n > 0 && presence == nullptr)) { set_error("ssrs_presence_allreduce: bad arguments"); return SSRS_ERR_INVALID; }
    if (comm == nullptr || comm->size <= 1) return SSRS_OK;        // one rank: nothing to add
    if (!comm->allreduce_u32) { set_error("ssrs_presence_allreduce: communicator has no all-reduce"); return SSRS_ERR_INVALID; }
    if (comm->allreduce_u32(comm->ctx, presence, n, stream) != 0) return SSRS_ERR_CUDA;
    return SSRS_OK;
}
