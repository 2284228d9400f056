// nccl_exchange.cu -- Exchange over NCCL (NVLink 5 / NVSwitch on an 8xB200 box).
//
// NCCL is resolved with dlopen at first use, so libfqd_b200.so has no link-time dependency
// on it and single-GPU users never load it.  In a process that already imported torch the
// soname resolves to torch's bundled NCCL; otherwise to the system libnccl.so.2.
#include <dlfcn.h>
#include <nccl.h>
#include <string.h>

#include <mutex>

#include "common.h"
#include "exchange.h"

namespace fqd {

namespace {

struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
};

NcclApi g_api;
std::mutex g_mu;

int load_api()
{
    std::lock_guard<std::mutex> lock(g_mu);
    if (g_api.handle) return FQD_OK;
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) {
        set_error("cannot load libnccl.so.2: %s", dlerror());
        return FQD_ERR_NCCL;
    }
#define SYM(field, name)                                                   \
    g_api.field = reinterpret_cast<decltype(g_api.field)>(dlsym(h, name)); \
    if (!g_api.field) { set_error("libnccl is missing %s", name); return FQD_ERR_NCCL; }
    SYM(GetUniqueId, "ncclGetUniqueId")
    SYM(CommInitRank, "ncclCommInitRank")
    SYM(CommDestroy, "ncclCommDestroy")
    SYM(AllGather, "ncclAllGather")
    SYM(AllReduce, "ncclAllReduce")
    SYM(Broadcast, "ncclBroadcast")
    SYM(Send, "ncclSend")
    SYM(Recv, "ncclRecv")
    SYM(GroupStart, "ncclGroupStart")
    SYM(GroupEnd, "ncclGroupEnd")
    SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
    g_api.handle = h;
    return FQD_OK;
}

#define FQD_NCCL(call)                                                                   \
    do {                                                                                 \
        ncclResult_t _r = (call);                                                        \
        if (_r != ncclSuccess) {                                                         \
            set_error("NCCL error %s in %s", g_api.GetErrorString(_r), #call);           \
            return FQD_ERR_NCCL;                                                         \
        }                                                                                \
    } while (0)

struct NcclExchange : Exchange {
    ncclComm_t comm = nullptr;
    void *scratch = nullptr;      // staging of the padded all-gather (grown on demand, kept)
    size_t scratch_cap = 0;
    ~NcclExchange() override
    {
        if (scratch) cudaFree(scratch);
        if (comm) g_api.CommDestroy(comm);
    }
    int allgather(const void *send, void *recv, size_t bytes, cudaStream_t s) override
    {
        FQD_NCCL(g_api.AllGather(send, recv, bytes, ncclUint8, comm, s));
        return FQD_OK;
    }
    int allgatherv(const void *send, void *recv, const size_t *off, const size_t *bytes,
                   cudaStream_t s) override
    {
        // Parts are hash-balanced, so pad them to the largest and use the real all-gather
        // (ring / NVLS tuned) instead of `world` broadcasts; then compact with local copies.
        size_t mx = 0, total = 0;
        for (int g = 0; g < world; g++) { mx = bytes[g] > mx ? bytes[g] : mx; total += bytes[g]; }
        if (!mx) return FQD_OK;
        mx = (mx + 255) & ~(size_t)255;
        if (scratch_cap < mx * (size_t)(world + 1)) {
            if (scratch) cudaFree(scratch);
            scratch = nullptr;
            scratch_cap = mx * (size_t)(world + 1) + (mx * (size_t)(world + 1)) / 4;
            FQD_CUDA(cudaMalloc(&scratch, scratch_cap));
        }
        char *stage_in = static_cast<char *>(scratch);
        char *stage_out = stage_in + mx;
        if (bytes[rank]) FQD_CUDA(cudaMemcpyAsync(stage_in, send, bytes[rank], cudaMemcpyDeviceToDevice, s));
        FQD_NCCL(g_api.AllGather(stage_in, stage_out, mx, ncclUint8, comm, s));
        for (int g = 0; g < world; g++)
            if (bytes[g])
                FQD_CUDA(cudaMemcpyAsync(static_cast<char *>(recv) + off[g], stage_out + (size_t)g * mx, bytes[g],
                                         cudaMemcpyDeviceToDevice, s));
        return FQD_OK;
    }
    int alltoallv(const void *send, const size_t *send_off, const size_t *send_bytes, void *recv,
                  const size_t *recv_off, const size_t *recv_bytes, cudaStream_t s) override
    {
        FQD_NCCL(g_api.GroupStart());
        for (int g = 0; g < world; g++) {
            if (send_bytes[g])
                FQD_NCCL(g_api.Send(static_cast<const char *>(send) + send_off[g], send_bytes[g], ncclUint8, g, comm, s));
            if (recv_bytes[g])
                FQD_NCCL(g_api.Recv(static_cast<char *>(recv) + recv_off[g], recv_bytes[g], ncclUint8, g, comm, s));
        }
        FQD_NCCL(g_api.GroupEnd());
        return FQD_OK;
    }
    int allreduce_max_u8(void *buf, size_t n, cudaStream_t s) override
    {
        FQD_NCCL(g_api.AllReduce(buf, buf, n, ncclUint8, ncclMax, comm, s));
        return FQD_OK;
    }
};

}  // namespace

int nccl_unique_id(uint8_t id[128])
{
    FQD_TRY(load_api());
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    ncclUniqueId uid;
    FQD_NCCL(g_api.GetUniqueId(&uid));
    memcpy(id, &uid, 128);
    return FQD_OK;
}

int nccl_exchange_create(int rank, int world, const uint8_t id[128], Exchange **out)
{
    *out = nullptr;
    FQD_TRY(load_api());
    ncclUniqueId uid;
    memcpy(&uid, id, 128);
    NcclExchange *ex = new NcclExchange();
    ex->rank = rank;
    ex->world = world;
    ncclResult_t r = g_api.CommInitRank(&ex->comm, world, uid, rank);
    if (r != ncclSuccess) {
        set_error("ncclCommInitRank failed: %s", g_api.GetErrorString(r));
        ex->comm = nullptr;
        delete ex;
        return FQD_ERR_NCCL;
    }
    *out = ex;
    return FQD_OK;
}

}  // namespace fqd
