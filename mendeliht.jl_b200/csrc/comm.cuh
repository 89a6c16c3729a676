#pragma once
#include "common.cuh"
#include <memory>
#include <stdlib.h>
#include <string.h>

constexpr int P2P_MAX_RANKS = 8;

struct ihtb_comm {
    void* comm = nullptr;   // ncclComm_t (NULL when nranks == 1)
    int rank = 0, nranks = 1, device = 0;
    int64_t n_collectives = 0;
    // peer-memory path (p2p.cu); p2p_local == NULL -> NCCL only
    bool p2p_tried = false;
    uint8_t* p2p_local = nullptr;
    std::vector<uint8_t*> p2p_peer;
    size_t p2p_slot_elems = 0;
    unsigned long long p2p_seq = 0;
    unsigned* p2p_counter = nullptr;
    int* p2p_err = nullptr;
};

namespace ihtb {
struct NcclUniqueId { char internal[128]; };
void comm_allreduce_sum_f64(ihtb_comm* c, double* d_buf, size_t count, cudaStream_t s);
void comm_allgather_i64(ihtb_comm* c, const int64_t* d_send, int64_t* d_recv, size_t count_per_rank, cudaStream_t s);

// what the peer-memory kernels see for the current sequence number
struct P2PView {
    double* push_slot[P2P_MAX_RANKS];                 // on rank r: slot[parity][my_rank]
    unsigned long long* push_flag[P2P_MAX_RANKS];     // on rank r: flag[parity][my_rank]
    const double* local_slot[P2P_MAX_RANKS];          // in my memory: slot[parity][r]
    const unsigned long long* local_flag;             // in my memory: flag[parity][0..nranks)
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

bool p2p_ready(const ihtb_comm* c, size_t n);
void p2p_setup(ihtb_comm* c, size_t n, cudaStream_t s, bool any_size = false);
void p2p_teardown(ihtb_comm* c);
P2PView p2p_view(ihtb_comm* c);
void p2p_push(ihtb_comm* c, const double* d_src, size_t n, cudaStream_t s);
void p2p_reduce(ihtb_comm* c, double* d_out, size_t n, cudaStream_t s);
bool p2p_failed(ihtb_comm* c);
// support.cu: fused producer (partial X[:,idx]*coef stored into every rank's slot)
void x_support_push(const ihtb_geno* g, const int64_t* d_idx, int64_t k, const double* d_coef, ihtb_comm* c,
                    cudaStream_t s);
}  // namespace ihtb
