#pragma once
#include "common.cuh"

namespace ihtb {

// Group structure of a doubly sparse fit (keywords J / k / group, reference src/fit.jl:64-68,
// project_group_sparse! src/utilities.jl:613-679).
struct GroupCtx {
    int64_t p = 0;
    int G = 0;                      // number of groups (ids 0..G-1 on the host, 1-based at the ABI)
    int J = 1;
    bool ks_vector = false;         // k given per group (the reference's `ks`)
    int64_t kcap = 0;               // scalar k the list capacities were sized for
    std::vector<int32_t> grp;       // [p_global] group of every SNP of the whole matrix (global column index)
    std::vector<int64_t> ks;        // [G] (ks_vector only)
    std::vector<int64_t> gsize, goff;   // members per group; offset of each group's candidate list (goff[G] = total)
    int64_t lcap = 0;               // longest candidate list
    DBuf<int64_t> d_order, d_gptr, d_goff, d_ks, d_gidx, d_oidx;
    DBuf<double> d_smax, d_gT, d_oval;   // d_gT = [T_L[G] | T_U[G] | overflow flag]
    DBuf<int32_t> d_chosen;
    DBuf<int64_t> d_Tall, d_blk, d_blkall;      // sharded fits: gathered bounds and (index, value) blocks
    HBuf<int64_t> h_Tall, h_blkall;
    HBuf<double> h_gT, h_oval;
    HBuf<int64_t> h_oidx;
    int chosen_cap = 0;

    int64_t k_of(int g, int64_t kscalar) const { return ks_vector ? ks[(size_t)g] : kscalar; }
    void build(int64_t p_, int64_t j0, int64_t p_global, const int32_t* group1, int J_, const int64_t* ks_,
               int64_t n_groups, int64_t kscalar, const double* h_sinv);
    void ensure_chosen(int n);
};

// Per group g, from the sweep's df (absolute error of entry j at most e_j = sinv_j * bound): every member that can be
// among the 2 k_g largest |df| of the group -> d_gidx (-1 padded), and bounds T_L <= T_g <= T_U on the sum of the k_g
// largest df^2 -> d_gT.  bound = bound_coef * (d_scal[1] + |d_scal[0]|) when d_scal != NULL, else host_bound.
void group_topk(GroupCtx& c, const double* d_dfa, const double* d_sinv, const double* d_scal, double bound_coef,
                double host_bound, int64_t kscalar, cudaStream_t s);
// copy the candidate lists of `n` chosen groups (ids in c.d_chosen) into c.d_oidx, lcap slots per group (-1 padded)
void group_take(GroupCtx& c, int n, cudaStream_t s);
void group_pack(const int64_t* d_oidx, const double* d_oval, int64_t slots, int64_t j0, int64_t* d_block, cudaStream_t s);

}  // namespace ihtb
