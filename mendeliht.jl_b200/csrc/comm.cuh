#pragma once
#include "common.cuh"
#include <memory>
#include <stdlib.h>

struct ihtb_comm {
    void* comm = nullptr;   // ncclComm_t (NULL when nranks == 1)
    int rank = 0, nranks = 1, device = 0;
    int64_t n_collectives = 0;
};

namespace ihtb {
struct NcclUniqueId { char internal[128]; };
void comm_allreduce_sum_f64(ihtb_comm* c, double* d_buf, size_t count, cudaStream_t s);
void comm_allgather_i64(ihtb_comm* c, const int64_t* d_send, int64_t* d_recv, size_t count_per_rank, cudaStream_t s);
}  // namespace ihtb
