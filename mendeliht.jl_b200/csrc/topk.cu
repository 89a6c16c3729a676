// Top-k projection support: candidate selection for project_k! (reference src/utilities.jl:553-559) over
// v_j = b0_j + eta * df_j, j = 1..p, with an error bound on df.
//
// The sweep's df_j may carry an absolute error of at most e_j = eta * sinv_j * bound (FAST mode: FP32 partial sums).
// With L_j = |v_j| - e_j and U_j = |v_j| + e_j, let tau be the k-th largest L.  The true k-th largest |v| is >= tau,
// so every member of the true top-k has U_j >= tau.  This file finds tau with a two-digit radix select on the
// order-preserving uint32 image of L (rounded down to FP32; the top 22 bits, two 11-bit digits, decide), then compacts
// {j : U_j >= tau} (U rounded up).  The fit
// re-scores those few columns exactly in FP64 and takes the top-k of the exact values, ties broken by lowest index,
// so the selected support does not depend on the sweep arithmetic.  Only integer atomics: deterministic.
#include "topk.cuh"
#include "comm.cuh"

namespace ihtb {

constexpr int TK_THREADS = 256;
constexpr int TK_BINS = 2048;

static inline int tk_grid(int64_t p) {
    int64_t b = ceil_div(p, TK_THREADS * 4);
    return (int)(b < 1 ? 1 : (b > 592 ? 592 : b));
}

__device__ __forceinline__ void hist_flush(const int* sh, int* g, int nb) {
    for (int b = threadIdx.x; b < nb; b += blockDim.x)
        if (sh[b]) atomicAdd(&g[b], sh[b]);
}

// pass 0: keys + histogram of the top 11 bits of keyL
__global__ void __launch_bounds__(TK_THREADS)
k_keys_hist0(int64_t p, const double* __restrict__ dfa, const double* __restrict__ b0d,
             const double* __restrict__ sinv, int64_t p_mod, const double* __restrict__ bounds, double eta,
             double bound, const double* __restrict__ scal, double bound_coef, const double* __restrict__ wt,
             uint32_t* __restrict__ keyL, uint32_t* __restrict__ keyU, int* __restrict__ hist,
             const double* __restrict__ l2, int* __restrict__ hist_other, TopkState* __restrict__ st,
             int64_t* __restrict__ cand, int cand_fill) {
    topk_first_kernel_housekeeping(hist_other, st, cand, cand_fill);
    // scal != NULL: the bound comes from the score kernel's sums still on the device:
    // bound = coef * (sum|r| + |sum r|) >= coef * ||r - mean(r)||_1 (no host round trip between sweep and selection)
    // l2 != NULL (PAIR sweeps): bound = coef * ||r - mean(r)||_2, and `sinv` is the handle's sgn array
    if (l2) bound = bound_coef * l2[0];
    else if (scal) bound = bound_coef * (scal[1] + fabs(scal[0]));
    __shared__ int sh[TK_BINS];
    for (int b = threadIdx.x; b < TK_BINS; b += blockDim.x) sh[b] = 0;
    __syncthreads();
    for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < p; j += (int64_t)gridDim.x * blockDim.x) {
        double v = (b0d ? b0d[j] : 0.0) + eta * dfa[j];
        // prior weights (src/utilities.jl:291-315): the projection ranks |v_j| * w_j, so magnitude and error scale by w_j
        const double w = wt ? wt[j] : 1.0;
        // blocked form (multivariate: entry j = t*p_mod + column): per-block bound, sinv of the column
        const double bj = bounds ? bounds[j / p_mod] : bound;
        uint32_t kl, ku;
        topk_make_keys(v, w, fabs(eta) * sinv[bounds ? (j % p_mod) : j] * bj * w, kl, ku);
        keyL[j] = kl; keyU[j] = ku;
        atomicAdd(&sh[kl >> 21], 1);
    }
    __syncthreads();
    hist_flush(sh, hist, TK_BINS);
}

// ---- fused digit picks -------------------------------------------------------------------------------------------
// Round 1 ran the select as reset / keys+hist0 / pick / hist / pick / hist / pick / compact = 8 dependent launches of a
// few microseconds each.  Every later pass now RE-DERIVES the digits of the earlier passes from their (finished,
// read-only) histograms inside each CTA -- a 2048-bin suffix scan per pass and CTA, a microsecond of redundant work --
// so the chain is keys+hist0 / hist1 / compact (a third digit pass existed until round 2b, see k_compact_fused).  Histograms are double-buffered between two consecutive selections
// (`set`): the first kernel of a selection clears the other set for the next one, nobody ever clears what is being read.
constexpr int TK_HIST_STRIDE = 3 * TK_BINS;        // ints per histogram set: hist0 | hist1 | hist2

// block-wide: digit of the bin (counted from the TOP) where the suffix count of `hist` crosses k; returns the bin and
// the remaining rank inside it through shared memory.  All TK_THREADS threads must call it.
__device__ __forceinline__ void block_pick(const int* __restrict__ hist, int nb, int k, int* sh_scan /*[32]*/,
                                           int* sh_out /*[2]*/) {
    constexpr int PER = TK_BINS / TK_THREADS;      // 8 reversed positions per thread
    const int t = threadIdx.x;
    int c[PER];
    int s = 0;
#pragma unroll
    for (int i = 0; i < PER; ++i) {
        const int r = t * PER + i;                 // reversed position: bin nb - 1 - r
        c[i] = (r < nb) ? hist[nb - 1 - r] : 0;
        s += c[i];
    }
    int incl = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int v = __shfl_up_sync(0xffffffffu, incl, o);
        if ((t & 31) >= o) incl += v;
    }
    if ((t & 31) == 31) sh_scan[t >> 5] = incl;
    __syncthreads();
    if (t < 32) {
        int v = (t < TK_THREADS / 32) ? sh_scan[t] : 0;
        int sc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int u = __shfl_up_sync(0xffffffffu, sc, o);
            if (t >= o) sc += u;
        }
        if (t < TK_THREADS / 32) sh_scan[t] = sc - v;          // exclusive prefix of the warp totals
    }
    __syncthreads();
    incl += sh_scan[t >> 5];
    int excl = incl - s;
    if (excl < k && k <= incl) {                                // exactly one thread: the crossing lies in its 8 bins
#pragma unroll
        for (int i = 0; i < PER; ++i) {
            if (excl < k && k <= excl + c[i]) { sh_out[0] = nb - 1 - (t * PER + i); sh_out[1] = k - excl; }
            excl += c[i];
        }
    }
    __syncthreads();
}

// pass 1 / 2: derive the digits of the earlier passes, then histogram the next digit among the keys that match
__global__ void __launch_bounds__(TK_THREADS)
k_hist_fused(int64_t p, const uint32_t* __restrict__ keyL, const int* __restrict__ hists /*this set*/, int pass, int kk,
             int* __restrict__ hist_out) {
    __shared__ int sh[TK_BINS];
    __shared__ int sh_scan[32];
    __shared__ int sh_out[2];
    block_pick(hists, TK_BINS, kk, sh_scan, sh_out);                           // digit 0: bits 31..21
    uint32_t prefix = (uint32_t)sh_out[0] << 21;
    int shift = 10, bits = 11;
    uint32_t himask = 0xFFE00000u;
    if (pass == 2) {
        const int k1 = sh_out[1];
        __syncthreads();
        block_pick(hists + TK_BINS, TK_BINS, k1, sh_scan, sh_out);             // digit 1: bits 20..10
        prefix |= (uint32_t)sh_out[0] << 10;
        shift = 0; bits = 10; himask = 0xFFFFFC00u;
    }
    const int nb = 1 << bits;
    for (int b = threadIdx.x; b < nb; b += blockDim.x) sh[b] = 0;
    __syncthreads();
    for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < p; j += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t k = keyL[j];
        if ((k & himask) == prefix) atomicAdd(&sh[(k >> shift) & (nb - 1)], 1);
    }
    __syncthreads();
    hist_flush(sh, hist_out, nb);
}

// last pass: derive both digits (tau = lower edge of the 22-bit bin of the k-th largest lower-bound key), compact {j : keyU_j >= tau}
__global__ void __launch_bounds__(TK_THREADS)
k_compact_fused(int64_t p, const uint32_t* __restrict__ keyU, const int* __restrict__ hists, int kk,
                TopkState* __restrict__ st, int64_t* __restrict__ cand, int cap) {
    __shared__ int sh_scan[32];
    __shared__ int sh_out[2];
    block_pick(hists, TK_BINS, kk, sh_scan, sh_out);
    uint32_t tau = (uint32_t)sh_out[0] << 21;
    int k1 = sh_out[1];
    __syncthreads();
    block_pick(hists + TK_BINS, TK_BINS, k1, sh_scan, sh_out);
    tau |= (uint32_t)sh_out[0] << 10;
    // Two digits (22 of the 32 key bits: sign, exponent, 13 mantissa bits) are enough: tau is the LOWER edge of the bin that
    // holds the k-th largest lower-bound key, so {keyU >= tau} is still a superset of the true top-k -- it admits the few
    // extra entries within 2^-13 of the threshold (re-scored exactly like the rest) and saves the third histogram pass.
    if (blockIdx.x == 0 && threadIdx.x == 0) { st->prefix = tau; st->k_rem = sh_out[1]; }
    for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < p; j += (int64_t)gridDim.x * blockDim.x) {
        if (keyU[j] >= tau) {
            int pos = atomicAdd(&st->count, 1);
            if (pos < cap) cand[pos] = j;
        }
    }
}

// dense b0 maintenance: zero old support entries, write new ones
__global__ void k_scatter(double* __restrict__ dst, const int64_t* __restrict__ idx, const double* __restrict__ val,
                          int64_t k, int zero_only) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t < k) dst[idx[t]] = zero_only ? 0.0 : val[t];
}

static void topk_run(TopkCtx& c, const double* d_dfa, const double* d_b0d, const double* d_sinv, int64_t p_mod,
                     const double* d_bounds, double eta, double bound, int64_t k, cudaStream_t s,
                     const double* d_scal = nullptr, double bound_coef = 0.0, const double* d_l2 = nullptr,
                     bool fill_cand = false) {
    int grid = tk_grid(c.p);
    c.fused_hist = nullptr;                 // a complete selection supersedes any first stage left pending
    const int64_t total = c.comm ? c.p_total : c.p;
    int kk = (int)(k < total ? k : total);
    int* hs = c.hist + (size_t)(c.set & 1) * TK_HIST_STRIDE;             // this selection's histograms (all zero)
    int* ho = c.hist + (size_t)((c.set & 1) ^ 1) * TK_HIST_STRIDE;       // cleared now for the next selection
    ++c.set;
    IHTB_LAUNCH(k_keys_hist0, grid, TK_THREADS, 0, s, c.p, d_dfa, d_b0d, d_sinv, p_mod, d_bounds, eta, bound, d_scal,
                bound_coef, c.wt, c.keyL, c.keyU, hs, d_l2, ho, c.st, c.cand, fill_cand ? c.cap : 0);
    if (c.comm) comm_allreduce_sum_i32(c.comm, hs, TK_BINS, s);
    IHTB_LAUNCH(k_hist_fused, grid, TK_THREADS, 0, s, c.p, c.keyL, hs, 1, kk, hs + TK_BINS);
    if (c.comm) comm_allreduce_sum_i32(c.comm, hs + TK_BINS, TK_BINS, s);
    IHTB_LAUNCH(k_compact_fused, grid, TK_THREADS, 0, s, c.p, c.keyU, hs, kk, c.st, c.cand, c.cap);
}

// The selection by |df| with its first stage (keys, first digit histogram, housekeeping) run by the sweep epilogue:
// begin() reserves the histogram set and describes the stage, finish() runs the second digit pass and the compaction.
TopkFuse topk_absdf_fuse_begin(TopkCtx& c, const double* d_sinv, const double* d_scal, double bound_coef) {
    static_assert(TOPK_BINS == TK_BINS, "one bin count");
    int* hs = c.hist + (size_t)(c.set & 1) * TK_HIST_STRIDE;
    int* ho = c.hist + (size_t)((c.set & 1) ^ 1) * TK_HIST_STRIDE;
    ++c.set;
    c.fused_hist = hs;
    TopkFuse f;
    f.keyL = c.keyL; f.keyU = c.keyU; f.hist = hs; f.hist_other = ho; f.st = c.st; f.cand = c.cand; f.cand_fill = c.cap;
    f.scale = d_sinv; f.scal = d_scal; f.bound_coef = bound_coef; f.wt = c.wt;
    return f;
}
void topk_candidates_absdf_finish(TopkCtx& c, int64_t k, cudaStream_t s) {
    IHTB_CHECK(c.fused_hist != nullptr && !c.comm, IHTB_EINVAL, "no fused selection stage pending");
    int* hs = c.fused_hist;
    c.fused_hist = nullptr;
    const int grid = tk_grid(c.p);
    const int kk = (int)(k < c.p ? k : c.p);
    IHTB_LAUNCH(k_hist_fused, grid, TK_THREADS, 0, s, c.p, c.keyL, hs, 1, kk, hs + TK_BINS);
    IHTB_LAUNCH(k_compact_fused, grid, TK_THREADS, 0, s, c.p, c.keyU, hs, kk, c.st, c.cand, c.cap);
}

void topk_candidates(TopkCtx& c, const double* d_dfa, const double* d_b0d, const double* d_sinv, double eta,
                     double bound, int64_t k, cudaStream_t s) {
    topk_run(c, d_dfa, d_b0d, d_sinv, c.p, nullptr, eta, bound, k, s);
}

// candidates by |df_j| alone (eta-independent): the top-k of |b0 + eta*df| always lies in supp(b0) plus the k largest
// |df_j| outside the support, whatever eta is, so ONE selection per sweep serves the gradient step and all of its
// backtracks.  The error bound is computed on the device from the score sums d_scal = [sum r, sum |r|, ...].
void topk_candidates_absdf(TopkCtx& c, const double* d_dfa, const double* d_sinv, const double* d_scal,
                           double bound_coef, int64_t k, cudaStream_t s, double host_bound, const double* d_l2) {
    // unused candidate slots stay -1 so that a gather launched over a fixed number of slots can skip them
    topk_run(c, d_dfa, nullptr, d_sinv, c.p, nullptr, 1.0, (d_scal || d_l2) ? 0.0 : host_bound, k, s, d_scal, bound_coef,
             d_l2, /*fill_cand=*/true);
}

__global__ void k_take(const double* __restrict__ src, const int64_t* __restrict__ idx, int64_t k,
                       double* __restrict__ dst) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t < k) dst[t] = src[idx[t]];
}
void take_values(const double* d_src, const int64_t* d_idx, int64_t k, double* d_dst, cudaStream_t s) {
    if (k) IHTB_LAUNCH(k_take, (unsigned)ceil_div(k, 128), 128, 0, s, d_src, d_idx, k, d_dst);
}

// entries e = t*p_mod + j (t-th right-hand side of column j): bound d_bounds[t], scale d_sinv[j]
void topk_candidates_blocked(TopkCtx& c, const double* d_dfa, const double* d_b0d, const double* d_sinv, int64_t p_mod,
                             const double* d_bounds, double eta, int64_t k, cudaStream_t s) {
    topk_run(c, d_dfa, d_b0d, d_sinv, p_mod, d_bounds, eta, 0.0, k, s);
}

// block = [count, 0, idx[capx], bits(val)[capx]] for the sharded candidate all-gather
__global__ void k_pack_candidates(int64_t* __restrict__ block, int64_t count, const int64_t* __restrict__ gidx,
                                  const double* __restrict__ vals, int capx) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t == 0) { block[0] = count; block[1] = 0; }
    if (t < count) {
        block[2 + t] = gidx[t];
        block[2 + capx + t] = __double_as_longlong(vals[t]);
    }
}
// Sharded fits, once per sweep: block = [n_cand, n_supp, idx[capx], bits(val)[capx]]; slots [0, capx/2) hold this rank's
// device-selected candidates (local index + j0, exact df from the gather over the first `glaunch` slots), slots
// [capx/2, capx) its part of the current support.
__global__ void k_pack_sweep(int64_t* __restrict__ block, const TopkState* __restrict__ st,
                             const int64_t* __restrict__ cand, int glaunch, const double* __restrict__ cand_vals,
                             const int64_t* __restrict__ supp, int nsupp, const double* __restrict__ supp_vals,
                             int64_t j0, int capx) {
    const int half = capx / 2;
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t == 0) { block[0] = st->count; block[1] = nsupp; }
    if (t < half) {
        const bool ok = t < st->count && t < glaunch;
        block[2 + t] = ok ? cand[t] + j0 : -1;
        block[2 + capx + t] = ok ? __double_as_longlong(cand_vals[t]) : 0;
    } else if (t < capx) {
        const int u = t - half;
        const bool ok = u < nsupp;
        block[2 + t] = ok ? supp[u] + j0 : -1;
        block[2 + capx + t] = ok ? __double_as_longlong(supp_vals[u]) : 0;
    }
}
void pack_sweep_candidates(int64_t* d_block, const TopkState* d_st, const int64_t* d_cand, int glaunch,
                           const double* d_cand_vals, const int64_t* d_supp, int nsupp, const double* d_supp_vals,
                           int64_t j0, int capx, cudaStream_t s) {
    IHTB_LAUNCH(k_pack_sweep, (unsigned)ceil_div(capx, 128), 128, 0, s, d_block, d_st, d_cand, glaunch, d_cand_vals,
                d_supp, nsupp, d_supp_vals, j0, capx);
}

void pack_candidates(int64_t* d_block, int64_t count, const int64_t* d_gidx, const double* d_vals, int capx,
                     cudaStream_t s) {
    IHTB_LAUNCH(k_pack_candidates, (unsigned)ceil_div(count > 0 ? count : 1, 128), 128, 0, s, d_block, count, d_gidx,
                d_vals, capx);
}

void scatter_dense(double* d_dst, const int64_t* d_idx, const double* d_val, int64_t k, int zero_only,
                   cudaStream_t s) {
    if (k == 0) return;
    IHTB_LAUNCH(k_scatter, (unsigned)ceil_div(k, 128), 128, 0, s, d_dst, d_idx, d_val, k, zero_only);
}

}  // namespace ihtb
