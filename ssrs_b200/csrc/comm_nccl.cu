// NCCL transport behind `ssrs_comm` (include/ssrs_b200.h): halo exchange, scalar all-reduce and all-gather for
// the row-sharded potential solve, and the presence-map all-reduce.  One process per GPU.
//
// Halo exchange over peer memory (SSRS_COMM_HALO=peer, when every rank can map its neighbours' staging blocks through
// CUDA IPC; the default stays grouped ncclSend/ncclRecv, see DESIGN §5 for the measurements): the sharded solve exchanges <= 100 KB halos about 26 times per
// BiCGStab iteration, and a grouped NCCL send/recv pair costs ~50 us of launch and proxy latency with almost nothing
// to move.  Here ONE kernel of four CTAs does an exchange: two CTAs store this rank's boundary ranges straight into
// the neighbours' staging buffers over NVLink and then release a sequence number there; two CTAs wait for the
// neighbours' sequence numbers in this rank's own block and copy the arrived ranges into the ghost entries.  Staging
// is double-buffered by the sequence number's parity; every exchange signals on every existing link, data or not, so
// a rank cannot run more than one exchange ahead of a neighbour (its write of exchange k + 1 follows its wait for
// exchange k, which the neighbour released after it had consumed exchange k - 1).
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
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
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
            SSRS_NCCL_SYM(Broadcast); SSRS_NCCL_SYM(AllGather); SSRS_NCCL_SYM(GetErrorString);
#undef SSRS_NCCL_SYM
            if (!a.GetUniqueId || !a.CommInitRank || !a.CommDestroy || !a.Send || !a.Recv || !a.GroupStart || !a.GroupEnd ||
                !a.AllReduce || !a.Broadcast)
                a.handle = nullptr;
        }
    }
    return a.handle ? &a : nullptr;
}

// ---- peer-memory halos -------------------------------------------------------------------------------------------
constexpr size_t kHaloMax = 4u << 20;                  // bytes per transfer (a fine-level row of 32767 cells is 256 KB)
constexpr size_t kFlagOff = 4 * kHaloMax;              // block: stage[from][parity] (4 x kHaloMax), then two 128-byte flag lines
constexpr size_t kBlockBytes = kFlagOff + 256;
constexpr long long kSpinLimit = 20000000000LL;        // clock64 ticks (~10 s) before a wait gives up and reports

struct HaloJob {            // one CTA's work
    const char* src;        // push: this rank's range; pull: this rank's staging
    char* dst;              // push: the neighbour's staging; pull: this rank's ghost range
    unsigned* flag;         // push: the neighbour's flag (written); pull: this rank's flag (awaited)
    long long nbytes;
    unsigned seq;
    int active;
};
struct HaloArgs { HaloJob job[4]; int* err; };         // 0, 1: push up / down; 2, 3: pull from up / down

__device__ __forceinline__ void copy_range(const char* src, char* dst, long long nbytes, bool src_is_staging) {
    // every range is a whole number of 4-byte elements; 16-byte accesses when both ends allow
    if ((((unsigned long long)src | (unsigned long long)dst | (unsigned long long)nbytes) & 15ull) == 0) {
        const uint4* s4 = reinterpret_cast<const uint4*>(src);
        uint4* d4 = reinterpret_cast<uint4*>(dst);
        for (long long i = threadIdx.x; i < (nbytes >> 4); i += blockDim.x) d4[i] = src_is_staging ? __ldcg(s4 + i) : s4[i];
    } else {
        const unsigned* s1 = reinterpret_cast<const unsigned*>(src);
        unsigned* d1 = reinterpret_cast<unsigned*>(dst);
        for (long long i = threadIdx.x; i < (nbytes >> 2); i += blockDim.x) d1[i] = src_is_staging ? __ldcg(s1 + i) : s1[i];
    }
}

__global__ void __launch_bounds__(512) halo_kernel(const HaloArgs A) {
    const HaloJob J = A.job[blockIdx.x];
    if (!J.active) return;
    if (blockIdx.x < 2) {
        // push: stores into the neighbour's memory, made visible system-wide by every storing thread, then the release
        copy_range(J.src, J.dst, J.nbytes, false);
        __threadfence_system();
        __syncthreads();
        if (threadIdx.x == 0) asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(J.flag), "r"(J.seq) : "memory");
    } else {
        if (threadIdx.x == 0) {
            const long long t0 = clock64();
            unsigned v;
            do {
                asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(J.flag) : "memory");
                if ((int)(v - J.seq) >= 0) break;
                if (clock64() - t0 > kSpinLimit) { *A.err = 1; break; }      // the neighbour never arrived: report, do not hang
            } while (true);
        }
        __syncthreads();
        copy_range(J.src, J.dst, J.nbytes, true);        // staging is read past L1 (ld.cg): the neighbour wrote it
    }
}

struct Ctx {
    ncclComm_t comm = nullptr;
    int rank = 0, size = 1;
    double* dscratch = nullptr;     // kScalars doubles on the device for small reductions (inner products; per-part integers of the setup)
    // peer-memory halos
    bool peer = false;
    char* block = nullptr;          // this rank's staging block (exported)
    char* nb[2] = {nullptr, nullptr};   // rank-1's and rank+1's blocks, mapped
    unsigned seq[2] = {0, 0};       // exchanges so far on the link to rank-1 / rank+1
    int* err_host = nullptr;        // mapped pinned flag a timed-out wait sets
    int* err_dev = nullptr;
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

// `from` index of a block's staging: 0 = data arriving from rank-1, 1 = from rank+1
inline char* stage_of(char* block, int from, unsigned seq) { return block + ((size_t)(2 * from) + (seq & 1u)) * kHaloMax; }
inline unsigned* flag_of(char* block, int from) { return reinterpret_cast<unsigned*>(block + kFlagOff + 128 * (size_t)from); }

int cb_exchange_peer(void* vctx, void* base, int64_t su_off, int64_t su_n, int64_t ru_off, int64_t ru_n,
                     int64_t sd_off, int64_t sd_n, int64_t rd_off, int64_t rd_n, void* stream) {
    Ctx* c = (Ctx*)vctx;
    char* b = (char*)base;
    const int64_t sizes[4] = {su_n, sd_n, ru_n, rd_n};
    for (int i = 0; i < 4; ++i)
        if (sizes[i] < 0 || (sizes[i] & 3) || (size_t)sizes[i] > kHaloMax) {
            set_error("ssrs_comm: halo of %lld bytes (peer staging holds %zu; unset SSRS_COMM_HALO for the NCCL path)", (long long)sizes[i], kHaloMax);
            return SSRS_ERR_INVALID;
        }
    if (*c->err_host) { set_error("ssrs_comm: a halo wait timed out earlier (a neighbouring rank stopped?)"); return SSRS_ERR_CUDA; }
    HaloArgs A;
    memset(&A, 0, sizeof(A));
    A.err = c->err_dev;
    const bool up = c->rank > 0, dn = c->rank + 1 < c->size;
    if (up) {
        const unsigned q = ++c->seq[0];
        // to rank-1: it sees me as its rank+1 (from = 1); from rank-1: my from = 0
        A.job[0] = HaloJob{b + su_off, stage_of(c->nb[0], 1, q), flag_of(c->nb[0], 1), su_n, q, 1};
        A.job[2] = HaloJob{stage_of(c->block, 0, q), b + ru_off, flag_of(c->block, 0), ru_n, q, 1};
    }
    if (dn) {
        const unsigned q = ++c->seq[1];
        A.job[1] = HaloJob{b + sd_off, stage_of(c->nb[1], 0, q), flag_of(c->nb[1], 0), sd_n, q, 1};
        A.job[3] = HaloJob{stage_of(c->block, 1, q), b + rd_off, flag_of(c->block, 1), rd_n, q, 1};
    }
    if (!up && !dn) return 0;
    halo_kernel<<<4, 512, 0, (cudaStream_t)stream>>>(A);
    SSRS_CUDA_TRY(cudaGetLastError());
    return 0;
}

// Maps the neighbours' staging blocks; collective.  Returns false (and leaves the NCCL path in place) when any rank
// cannot: all ranks must use the same transport.
bool setup_peer_halos(Ctx* c, NcclApi* a) {
    if (c->size < 2 || !a->AllGather) return false;
    const char* mode = getenv("SSRS_COMM_HALO");
    int want = mode && strcmp(mode, "peer") == 0;          // opt-in: measured equal to NCCL on 2 GPUs
    cudaIpcMemHandle_t mine;
    memset(&mine, 0, sizeof(mine));
    char* all_d = nullptr;
    cudaIpcMemHandle_t* all_h = (cudaIpcMemHandle_t*)malloc(sizeof(cudaIpcMemHandle_t) * (size_t)c->size);
    int ok = want;
    if (ok && cudaMalloc(&c->block, kBlockBytes) != cudaSuccess) ok = 0;
    if (ok && cudaMemset(c->block, 0, kBlockBytes) != cudaSuccess) ok = 0;
    if (ok && cudaIpcGetMemHandle(&mine, c->block) != cudaSuccess) ok = 0;
    if (ok && cudaHostAlloc((void**)&c->err_host, sizeof(int), cudaHostAllocMapped) != cudaSuccess) ok = 0;
    if (ok) { *c->err_host = 0; if (cudaHostGetDevicePointer((void**)&c->err_dev, c->err_host, 0) != cudaSuccess) ok = 0; }
    // the handles of all ranks (a rank that failed contributes zeros; the agreement below settles it)
    bool gathered = cudaMalloc(&all_d, sizeof(cudaIpcMemHandle_t) * (size_t)c->size) == cudaSuccess &&
                    cudaMemcpy(all_d + sizeof(mine) * (size_t)c->rank, &mine, sizeof(mine), cudaMemcpyHostToDevice) == cudaSuccess &&
                    a->AllGather(all_d + sizeof(mine) * (size_t)c->rank, all_d, sizeof(mine), ncclChar, c->comm, 0) == ncclSuccess &&
                    cudaMemcpy(all_h, all_d, sizeof(mine) * (size_t)c->size, cudaMemcpyDeviceToHost) == cudaSuccess;
    if (!gathered) ok = 0;
    if (ok && c->rank > 0 && cudaIpcOpenMemHandle((void**)&c->nb[0], all_h[c->rank - 1], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) ok = 0;
    if (ok && c->rank + 1 < c->size && cudaIpcOpenMemHandle((void**)&c->nb[1], all_h[c->rank + 1], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) ok = 0;
    cudaGetLastError();                                   // a failed attempt must not poison later calls
    // agreement: the minimum over the ranks
    double v = (double)ok;
    bool agreed = cudaMemcpy(c->dscratch, &v, sizeof(v), cudaMemcpyHostToDevice) == cudaSuccess &&
                  a->AllReduce(c->dscratch, c->dscratch, 1, ncclDouble, ncclMin, c->comm, 0) == ncclSuccess &&
                  cudaMemcpy(&v, c->dscratch, sizeof(v), cudaMemcpyDeviceToHost) == cudaSuccess;
    if (all_d) cudaFree(all_d);
    free(all_h);
    if (!(agreed && v == 1.0)) {
        for (int i = 0; i < 2; ++i) if (c->nb[i]) { cudaIpcCloseMemHandle(c->nb[i]); c->nb[i] = nullptr; }
        if (c->block) { cudaFree(c->block); c->block = nullptr; }
        if (c->err_host) { cudaFreeHost(c->err_host); c->err_host = nullptr; }
        cudaGetLastError();
        return false;
    }
    return true;
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
    if (c->peer && *c->err_host) { set_error("ssrs_comm: a halo wait timed out (a neighbouring rank stopped?)"); return SSRS_ERR_CUDA; }
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
    c->peer = setup_peer_halos(c, a);
    m->exchange = c->peer ? cb_exchange_peer : cb_exchange; m->allreduce_sum = cb_allreduce_sum; m->allgather = cb_allgather; m->allreduce_u32 = cb_allreduce_u32;
    *out = m;
    return SSRS_OK;
}

extern "C" int ssrs_comm_destroy(ssrs_comm* m) {
    if (!m) return SSRS_OK;
    NcclApi* a = api();
    Ctx* c = (Ctx*)m->ctx;
    if (c) {
        if (c->peer && a && c->comm) {
            // nobody unmaps or frees a block a neighbour may still be storing into: two collective fences
            cudaDeviceSynchronize();
            a->AllReduce(c->dscratch, c->dscratch, 1, ncclDouble, ncclSum, c->comm, 0); cudaDeviceSynchronize();
            for (int i = 0; i < 2; ++i) if (c->nb[i]) cudaIpcCloseMemHandle(c->nb[i]);
            a->AllReduce(c->dscratch, c->dscratch, 1, ncclDouble, ncclSum, c->comm, 0); cudaDeviceSynchronize();
            if (c->block) cudaFree(c->block);
            if (c->err_host) cudaFreeHost(c->err_host);
        }
        if (c->dscratch) cudaFree(c->dscratch);
        if (a && c->comm) a->CommDestroy(c->comm);
        delete c;
    }
    delete m;
    return SSRS_OK;
}

extern "C" int ssrs_comm_halo_mode(const ssrs_comm* m) {
    // 1: halos over peer memory (one kernel per exchange); 0: grouped ncclSend/ncclRecv; -1: not an NCCL communicator
    if (!m || !m->ctx || (m->exchange != cb_exchange && m->exchange != cb_exchange_peer)) return -1;
    return ((Ctx*)m->ctx)->peer ? 1 : 0;
}

extern "C" int ssrs_presence_allreduce(uint32_t* presence, int64_t n, const ssrs_comm* comm, void* stream) {
    if (n < 0 || (n > 0 && presence == nullptr)) { set_error("ssrs_presence_allreduce: bad arguments"); return SSRS_ERR_INVALID; }
    if (comm == nullptr || comm->size <= 1) return SSRS_OK;        // one rank: nothing to add
    if (!comm->allreduce_u32) { set_error("ssrs_presence_allreduce: communicator has no all-reduce"); return SSRS_ERR_INVALID; }
    if (comm->allreduce_u32(comm->ctx, presence, n, stream) != 0) return SSRS_ERR_CUDA;
    return SSRS_OK;
}
