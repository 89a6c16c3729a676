// NCCL plumbing for SNP-sharded fits (one process per GPU; no reference equivalent).  NCCL carries the
// rendezvous (IPC handles of the peer-memory region) and is the fallback when peer mapping is unavailable; the
// collectives of the IHT loop themselves run over peer memory (p2p.cu).
// libnccl is loaded at run time (dlopen) so that libihtb200.so itself has no link-time NCCL dependency and loads on
// machines without it; the torch-bundled libnccl.so.2 is reused when the host process already imported torch.
#include "comm.cuh"
#include <dlfcn.h>
#include <string.h>

namespace ihtb {

namespace {
struct NcclApi {
    void* handle = nullptr;
    int (*GetUniqueId)(NcclUniqueId*) = nullptr;
    int (*CommInitRank)(void**, int, NcclUniqueId, int) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};
NcclApi g_nccl;

void nccl_load(const char* path) {
    if (g_nccl.handle) return;
    const char* cands[] = {path, getenv("IHTB_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    void* h = nullptr;
    std::string tried;
    for (const char* c : cands) {
        if (!c || !*c) continue;
        h = dlopen(c, RTLD_NOW | RTLD_GLOBAL);
        if (h) break;
        tried += std::string(c) + " ";
    }
    IHTB_CHECK(h, IHTB_ECUDA, "cannot load NCCL (tried: " + tried + ")");
    auto sym = [&](const char* n) {
        void* s = dlsym(h, n);
        IHTB_CHECK(s, IHTB_ECUDA, std::string("NCCL symbol missing: ") + n);
        return s;
    };
    g_nccl.GetUniqueId = reinterpret_cast<decltype(g_nccl.GetUniqueId)>(sym("ncclGetUniqueId"));
    g_nccl.CommInitRank = reinterpret_cast<decltype(g_nccl.CommInitRank)>(sym("ncclCommInitRank"));
    g_nccl.CommDestroy = reinterpret_cast<decltype(g_nccl.CommDestroy)>(sym("ncclCommDestroy"));
    g_nccl.AllReduce = reinterpret_cast<decltype(g_nccl.AllReduce)>(sym("ncclAllReduce"));
    g_nccl.AllGather = reinterpret_cast<decltype(g_nccl.AllGather)>(sym("ncclAllGather"));
    g_nccl.GetErrorString = reinterpret_cast<decltype(g_nccl.GetErrorString)>(sym("ncclGetErrorString"));
    g_nccl.handle = h;
}

void nccl_check(int rc, const char* what) {
    if (rc != 0)
        throw Error(IHTB_ECUDA, std::string(what) + ": " + (g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "?"));
}
}  // namespace

constexpr int kNcclSum = 0, kNcclInt32 = 2, kNcclInt64 = 4, kNcclFloat64 = 8;

void nccl_allreduce_sum_f64(ihtb_comm* c, double* d_buf, size_t count, cudaStream_t s) {
    if (!c || c->nranks == 1 || count == 0) return;
    IHTB_CHECK(c->comm, IHTB_ECUDA, "this communicator has no NCCL backend and its peer memory is not mapped");
    nccl_check(g_nccl.AllReduce(d_buf, d_buf, count, kNcclFloat64, kNcclSum, c->comm, s), "ncclAllReduce");
    ++c->n_nccl_calls;
}

void nccl_allreduce_sum_i32(ihtb_comm* c, int* d_buf, size_t count, cudaStream_t s) {
    if (!c || c->nranks == 1 || count == 0) return;
    IHTB_CHECK(c->comm, IHTB_ECUDA, "this communicator has no NCCL backend and its peer memory is not mapped");
    nccl_check(g_nccl.AllReduce(d_buf, d_buf, count, kNcclInt32, kNcclSum, c->comm, s), "ncclAllReduce");
    ++c->n_nccl_calls;
}

void nccl_allgather_i64(ihtb_comm* c, const int64_t* d_send, int64_t* d_recv, size_t count_per_rank,
                        cudaStream_t s) {
    if (!c || c->nranks == 1) {
        if (d_send != d_recv)
            IHTB_CUDA(cudaMemcpyAsync(d_recv, d_send, count_per_rank * sizeof(int64_t), cudaMemcpyDeviceToDevice, s));
        return;
    }
    IHTB_CHECK(c->comm, IHTB_ECUDA, "this communicator has no NCCL backend and its peer memory is not mapped");
    nccl_check(g_nccl.AllGather(d_send, d_recv, count_per_rank, kNcclInt64, c->comm, s), "ncclAllGather");
    ++c->n_nccl_calls;
}

}  // namespace ihtb

using namespace ihtb;

extern "C" {

int32_t ihtb_comm_unique_id(const char* nccl_lib_path, uint8_t* out128) {
    return guard([&] {
        IHTB_CHECK(out128, IHTB_EINVAL, "NULL argument");
        nccl_load(nccl_lib_path);
        NcclUniqueId id;
        nccl_check(g_nccl.GetUniqueId(&id), "ncclGetUniqueId");
        memcpy(out128, id.internal, 128);
    });
}

int32_t ihtb_comm_create(const char* nccl_lib_path, const uint8_t* id128, int32_t rank, int32_t nranks,
                         ihtb_comm** out) {
    return guard([&] {
        IHTB_CHECK(id128 && out && nranks >= 1 && rank >= 0 && rank < nranks, IHTB_EINVAL, "bad argument");
        std::unique_ptr<ihtb_comm> c(new ihtb_comm());
        c->rank = rank; c->nranks = nranks;
        IHTB_CUDA(cudaGetDevice(&c->device));
        if (nranks > 1) {
            nccl_load(nccl_lib_path);
            NcclUniqueId id;
            memcpy(id.internal, id128, 128);
            nccl_check(g_nccl.CommInitRank(&c->comm, nranks, id, rank), "ncclCommInitRank");
        }
        *out = c.release();
    });
}

// counters since creation: collectives served by the library's peer-memory kernels, and NCCL calls (the rendezvous of
// the peer mappings, or every collective when peer mapping is unavailable)
int32_t ihtb_comm_stats(const ihtb_comm* c, int64_t* peer_memory_collectives, int64_t* nccl_calls) {
    return guard([&] {
        IHTB_CHECK(c, IHTB_EINVAL, "NULL argument");
        if (peer_memory_collectives) *peer_memory_collectives = c->n_collectives;
        if (nccl_calls) *nccl_calls = c->n_nccl_calls;
    });
}

int32_t ihtb_comm_destroy(ihtb_comm* c) {
    return guard([&] {
        if (!c) return;
        p2p_teardown(c);
        if (c->comm) g_nccl.CommDestroy(c->comm);
        delete c;
    });
}

}  // extern "C"
