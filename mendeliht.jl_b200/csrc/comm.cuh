// Communicator of SNP-sharded fits (no reference equivalent).  Two ways to form one:
//   * one process per GPU (torchrun): ranks rendezvous through NCCL (comm.cu); peer memory is mapped with CUDA IPC;
//   * one process driving several GPUs (ihtb_mgeno, multi.cu): ranks are host threads of this process that share a
//     LocalGroup; peer memory is mapped with cudaDeviceEnablePeerAccess.  No NCCL involved.
// Either way the collectives of the IHT loop run over a symmetric peer-memory region (p2p.cu); NCCL is only the
// fallback when peer mapping is unavailable in the multi-process case.
#pragma once
#include "common.cuh"
#include <condition_variable>
#include <memory>
#include <mutex>
#include <stdlib.h>
#include <string.h>

constexpr int P2P_MAX_RANKS = 8;
constexpr int P2P_PAR = 4;            // sequence parities: a slot is reused every fourth operation of its kind
constexpr int P2P_KINDS = 3;          // flag sets: 0 push-all / partial ready, 1 reduced slice ready, 2 gather block ready

namespace ihtb {
// ranks of one process (host threads): a reusable barrier and a pointer exchange board
struct LocalGroup {
    int nranks = 0;
    std::mutex mu;
    std::condition_variable cv;
    int waiting = 0;
    unsigned long long generation = 0;
    bool failed = false;                       // a rank threw: everyone leaves the barriers with an error
    void* board[P2P_MAX_RANKS] = {};
    int devices[P2P_MAX_RANKS] = {};
    void barrier();                            // throws on every rank once any rank called fail()
    void fail();
};
}  // namespace ihtb

struct ihtb_comm {
    void* comm = nullptr;   // ncclComm_t (NULL when nranks == 1 or for in-process groups)
    int rank = 0, nranks = 1, device = 0;
    int64_t n_collectives = 0;      // collectives served by the peer-memory kernels (p2p.cu)
    int64_t n_nccl_calls = 0;       // ncclAllReduce / ncclAllGather calls (rendezvous, or fallback without peer mapping)
    std::shared_ptr<ihtb::LocalGroup> local;      // in-process group (multi.cu); null for NCCL communicators
    // symmetric peer-memory region (p2p.cu); sym_local == NULL -> NCCL only
    bool p2p_tried = false;
    bool ipc_mapped = false;
    uint8_t* sym_local = nullptr;
    std::vector<uint8_t*> sym_peer;
    size_t pa_cap = 0;      // elements per push-all slot (small vectors: every rank stores to every rank)
    size_t red_cap = 0;     // elements of the two-phase all-reduce areas (partial, result)
    size_t gat_cap = 0;     // int64 per rank block of the all-gather area
    size_t off_pa = 0, off_partial = 0, off_result = 0, off_gather = 0, sym_bytes = 0;
    unsigned long long seq[P2P_KINDS] = {0, 0, 0};
    unsigned* p2p_counter = nullptr;
    int* p2p_err = nullptr;
};

namespace ihtb {
struct NcclUniqueId { char internal[128]; };
// generic collectives: peer memory when mapped (any size, chunked), NCCL otherwise
void comm_allreduce_sum_f64(ihtb_comm* c, double* d_buf, size_t count, cudaStream_t s);
void comm_allgather_i64(ihtb_comm* c, const int64_t* d_send, int64_t* d_recv, size_t count_per_rank, cudaStream_t s);
void comm_allreduce_sum_i32(ihtb_comm* c, int* d_buf, size_t count, cudaStream_t s);      // counts (exact)
void nccl_allreduce_sum_i32(ihtb_comm* c, int* d_buf, size_t count, cudaStream_t s);
void nccl_allreduce_sum_f64(ihtb_comm* c, double* d_buf, size_t count, cudaStream_t s);
void nccl_allgather_i64(ihtb_comm* c, const int64_t* d_send, int64_t* d_recv, size_t count_per_rank, cudaStream_t s);

// what the push-all producer kernels see for the current sequence number
struct P2PView {
    double* push_slot[P2P_MAX_RANKS];                 // on rank r: slot[parity][my_rank]
    unsigned long long* push_flag[P2P_MAX_RANKS];     // on rank r: flag[kind][parity][my_rank]
    const double* local_slot[P2P_MAX_RANKS];          // in my memory: slot[parity][r]
    const unsigned long long* local_flag;             // in my memory: flag[kind][parity][0..nranks)
    unsigned* counter;                                // CTAs finished (producer kernels)
    int* err;
    int nranks, rank;
};

// last step of a producer kernel: make this CTA's peer stores visible, and let the last CTA publish the flags
__device__ __forceinline__ void p2p_publish(const P2PView& v, unsigned long long seq) {
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned done = atomicAdd(v.counter, 1u);
        if (done == gridDim.x - 1) {
            // every CTA fenced its data before its counter increment; one more fence orders this thread's observation
            // of the counter before the flags, which then go out back to back (no fence per peer)
            *v.counter = 0u;
            __threadfence_system();
            for (int r = 0; r < v.nranks; ++r)
                asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(v.push_flag[r]), "l"(seq) : "memory");
        }
    }
}

// setup is collective: every rank calls it with the same arguments.  red_elems / gat_elems are the largest all-reduce /
// all-gather block the caller wants to run without chunking.
void p2p_setup(ihtb_comm* c, size_t red_elems, size_t gat_elems, cudaStream_t s);
void p2p_teardown(ihtb_comm* c);
bool p2p_mapped(const ihtb_comm* c);
bool p2p_pushall_ok(const ihtb_comm* c, size_t count);      // small vector: fused push-all path
bool p2p_failed(ihtb_comm* c);
// push-all all-reduce (count <= pa_cap): a producer stores its partial into every rank, one local reduce
P2PView p2p_view(ihtb_comm* c);
void p2p_push(ihtb_comm* c, const double* d_src, size_t count, cudaStream_t s);
void p2p_reduce(ihtb_comm* c, double* d_out, size_t count, cudaStream_t s);
// two-phase all-reduce (count <= red_cap): the producer writes its partial to p2p_partial_ptr(), then every rank
// reduces one slice from all peers and stores it into every rank's result area
double* p2p_partial_ptr(ihtb_comm* c);
void p2p_allreduce_2phase(ihtb_comm* c, size_t count, double* d_out, cudaStream_t s);
// all-gather of count <= gat_cap int64 per rank into d_recv[nranks][count]
void p2p_allgather(ihtb_comm* c, const int64_t* d_send, size_t count, int64_t* d_recv, cudaStream_t s);
// support.cu: fused producer (partial X[:,idx]*coef stored into every rank's slot)
void x_support_push(const ihtb_geno* g, const int64_t* d_idx, int64_t k, const double* d_coef, ihtb_comm* c,
                    cudaStream_t s);
}  // namespace ihtb
