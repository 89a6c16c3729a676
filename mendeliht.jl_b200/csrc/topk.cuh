#pragma once
#include "common.cuh"

struct ihtb_comm;

namespace ihtb {

struct TopkState {
    uint32_t prefix;   // after the 3 passes: tau = k-th largest lower-bound key
    int32_t k_rem;
    int32_t count;     // candidates found (may exceed cap)
    int32_t pad;
};

struct TopkCtx {
    int64_t p;
    uint32_t* keyL;
    uint32_t* keyU;
    int* hist;         // 2 sets x 3 x 2048 ints, zero at creation (double-buffered between selections, topk.cu)
    TopkState* st;
    int64_t* cand;
    int cap;
    const double* wt = nullptr;   // optional prior weights [p]: keys rank |v_j| * wt_j
    unsigned set = 0;             // histogram set of the next selection
    // Column-sharded selection over ALL shards' entries (mvfit.cu): the three digit histograms are all-reduced between
    // the passes, so every rank derives the same tau (the k-th largest lower-bound key of the whole problem) and compacts
    // only its own entries that can reach it.  p_total = entries over all shards.
    ihtb_comm* comm = nullptr;
    int64_t p_total = 0;
};

void topk_candidates(TopkCtx& c, const double* d_dfa, const double* d_b0d, const double* d_sinv, double eta,
                     double bound, int64_t k, cudaStream_t s);
void topk_candidates_blocked(TopkCtx& c, const double* d_dfa, const double* d_b0d, const double* d_sinv, int64_t p_mod,
                             const double* d_bounds, double eta, int64_t k, cudaStream_t s);
// d_scal != NULL: bound = bound_coef * (d_scal[1] + |d_scal[0]|) on the device; else host_bound is used
void topk_candidates_absdf(TopkCtx& c, const double* d_dfa, const double* d_sinv, const double* d_scal,
                           double bound_coef, int64_t k, cudaStream_t s, double host_bound = 0.0,
                           const double* d_l2 = nullptr);
void take_values(const double* d_src, const int64_t* d_idx, int64_t k, double* d_dst, cudaStream_t s);
void pack_sweep_candidates(int64_t* d_block, const TopkState* d_st, const int64_t* d_cand, int glaunch,
                           const double* d_cand_vals, const int64_t* d_supp, int nsupp, const double* d_supp_vals,
                           int64_t j0, int capx, cudaStream_t s);
void pack_candidates(int64_t* d_block, int64_t count, const int64_t* d_gidx, const double* d_vals, int capx,
                     cudaStream_t s);
void scatter_dense(double* d_dst, const int64_t* d_idx, const double* d_val, int64_t k, int zero_only, cudaStream_t s);

}  // namespace ihtb
