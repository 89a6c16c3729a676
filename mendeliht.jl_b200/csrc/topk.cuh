#pragma once
#include "common.cuh"

struct ihtb_comm;

namespace ihtb {

struct TopkState {
    uint32_t prefix;   // after the passes: tau = lower edge of the 22-bit bin of the k-th largest lower-bound key
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
    int* fused_hist = nullptr;    // set reserved by topk_absdf_fuse_begin, consumed by topk_candidates_absdf_finish
    // Column-sharded selection over ALL shards' entries (mvfit.cu): the digit histograms are all-reduced between
    // the passes, so every rank derives the same tau (the k-th largest lower-bound key of the whole problem) and compacts
    // only its own entries that can reach it.  p_total = entries over all shards.
    ihtb_comm* comm = nullptr;
    int64_t p_total = 0;
};

// First stage of a |df|-selection handed to the kernel that PRODUCES df (the sweep epilogue, sweep.cu): it computes the
// keys of its own columns and the first digit histogram, so the selection starts at its second pass.  keyL == NULL: unused.
struct TopkFuse {
    uint32_t* keyL = nullptr;
    uint32_t* keyU = nullptr;
    int* hist = nullptr;          // this selection's histogram set
    int* hist_other = nullptr;    // the other set, cleared for the next selection
    TopkState* st = nullptr;
    int64_t* cand = nullptr;
    int cand_fill = 0;
    const double* scale = nullptr;   // per-column scale of the error bound (sinv)
    const double* scal = nullptr;    // score sums on the device: bound = coef * (scal[1] + |scal[0]|)
    double bound_coef = 0.0;
    const double* wt = nullptr;
};
constexpr int TOPK_BINS = 2048;

// housekeeping of a selection's first kernel: the other histogram set is cleared for the NEXT selection, the candidate
// count restarts, unused candidate slots read -1 (a gather launched over a fixed number of slots skips them)
__device__ __forceinline__ void topk_first_kernel_housekeeping(int* hist_other, TopkState* st, int64_t* cand, int cand_fill) {
    if (blockIdx.x == 0) {
        for (int b = threadIdx.x; b < 3 * TOPK_BINS; b += blockDim.x) hist_other[b] = 0;
        if (threadIdx.x == 0) { st->prefix = 0; st->k_rem = 0; st->count = 0; st->pad = 0; }
    }
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < cand_fill; i += (int64_t)gridDim.x * blockDim.x)
        cand[i] = -1;
}
// order-preserving uint32 images of the lower / upper bound of |v| * w given the absolute error e of v * w
__device__ __forceinline__ void topk_make_keys(double v, double w, double e_abs, uint32_t& kl, uint32_t& ku) {
    const double a = fabs(v) * w;
    const double e = e_abs + a * 4e-16;
    double lo = a - e, up = a + e;
    if (!(lo > 0.0)) lo = 0.0;          // also maps NaN to 0
    if (!(up >= 0.0)) up = INFINITY;    // NaN: always a candidate
    kl = __float_as_uint(__double2float_rd(lo));
    ku = __float_as_uint(__double2float_ru(up));
}

// host side of the fused first stage: reserves this selection's histogram set; topk_candidates_absdf_finish runs the rest
TopkFuse topk_absdf_fuse_begin(TopkCtx& c, const double* d_sinv, const double* d_scal, double bound_coef);
void topk_candidates_absdf_finish(TopkCtx& c, int64_t k, cudaStream_t s);

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
