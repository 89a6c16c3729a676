// Top-k projection support: candidate selection for project_k! (reference src/utilities.jl:553-559) over
// v_j = b0_j + eta * df_j, j = 1..p, with an error bound on df.
//
// The sweep's df_j may carry an absolute error of at most e_j = eta * sinv_j * bound (FAST mode: FP32 partial sums).
// With L_j = |v_j| - e_j and U_j = |v_j| + e_j, let tau be the k-th largest L.  The true k-th largest |v| is >= tau,
// so every member of the true top-k has U_j >= tau.  This file finds tau with a 3-pass radix select on the
// order-preserving uint32 image of L (rounded down to FP32), then compacts {j : U_j >= tau} (U rounded up).  The fit
// re-scores those few columns exactly in FP64 and takes the top-k of the exact values, ties broken by lowest index,
// so the selected support does not depend on the sweep arithmetic.  Only integer atomics: deterministic.
#include "topk.cuh"

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
             const double* __restrict__ l2) {
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
        double a = fabs(v) * w;
        // blocked form (multivariate: entry j = t*p_mod + column): per-block bound, sinv of the column
        const double bj = bounds ? bounds[j / p_mod] : bound;
        double e = fabs(eta) * sinv[bounds ? (j % p_mod) : j] * bj * w + a * 4e-16;
        double lo = a - e, up = a + e;
        if (!(lo > 0.0)) lo = 0.0;          // also maps NaN to 0
        if (!(up >= 0.0)) up = INFINITY;    // NaN: always a candidate
        uint32_t kl = __float_as_uint(__double2float_rd(lo));
        uint32_t ku = __float_as_uint(__double2float_ru(up));
        keyL[j] = kl; keyU[j] = ku;
        atomicAdd(&sh[kl >> 21], 1);
    }
    __syncthreads();
    hist_flush(sh, hist, TK_BINS);
}

// later passes: histogram of `bits` bits at `shift` among keys whose higher bits equal st->prefix
__global__ void __launch_bounds__(TK_THREADS)
k_hist(int64_t p, const uint32_t* __restrict__ keyL, const TopkState* __restrict__ st, int shift, int bits,
       int* __restrict__ hist) {
    __shared__ int sh[TK_BINS];
    const int nb = 1 << bits;
    for (int b = threadIdx.x; b < nb; b += blockDim.x) sh[b] = 0;
    __syncthreads();
    const uint32_t prefix = st->prefix;
    const uint32_t himask = ~((1u << (shift + bits)) - 1u);   // bits above this digit
    for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < p; j += (int64_t)gridDim.x * blockDim.x) {
        uint32_t k = keyL[j];
        if ((k & himask) == (prefix & himask)) atomicAdd(&sh[(k >> shift) & (nb - 1)], 1);
    }
    __syncthreads();
    hist_flush(sh, hist, nb);
}

// one block of 1024 threads: find the digit whose suffix count crosses k_rem, then clear the histogram
__global__ void __launch_bounds__(1024)
k_pick(int* __restrict__ hist, TopkState* __restrict__ st, int shift, int bits) {
    __shared__ int warp_tot[32];
    const int nb = 1 << bits;
    const int t = threadIdx.x;
    // reversed order: position r <-> bin nb-1-r, two positions per thread
    const int r0 = 2 * t, r1 = 2 * t + 1;
    const int c0 = (r0 < nb) ? hist[nb - 1 - r0] : 0;
    const int c1 = (r1 < nb) ? hist[nb - 1 - r1] : 0;
    int s = c0 + c1;
    int incl = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int v = __shfl_up_sync(0xffffffffu, incl, o);
        if ((t & 31) >= o) incl += v;
    }
    if ((t & 31) == 31) warp_tot[t >> 5] = incl;
    __syncthreads();
    if (t < 32) {
        int v = warp_tot[t];
        int sc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int u = __shfl_up_sync(0xffffffffu, sc, o);
            if (t >= o) sc += u;
        }
        warp_tot[t] = sc - v;   // exclusive prefix of warp totals
    }
    __syncthreads();
    incl += warp_tot[t >> 5];
    const int excl = incl - s;
    const int k = st->k_rem;
    __syncthreads();
    if (excl < k && k <= incl) {
        int bin, krem;
        if (excl + c0 >= k) { bin = nb - 1 - r0; krem = k - excl; }
        else { bin = nb - 1 - r1; krem = k - excl - c0; }
        st->prefix |= (uint32_t)bin << shift;
        st->k_rem = krem;
    }
    for (int b = t; b < TK_BINS; b += blockDim.x) hist[b] = 0;
}

__global__ void __launch_bounds__(TK_THREADS)
k_compact(int64_t p, const uint32_t* __restrict__ keyU, TopkState* __restrict__ st, int64_t* __restrict__ cand,
          int cap) {
    const uint32_t tau = st->prefix;
    for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < p; j += (int64_t)gridDim.x * blockDim.x) {
        if (keyU[j] >= tau) {
            int pos = atomicAdd(&st->count, 1);
            if (pos < cap) cand[pos] = j;
        }
    }
}

__global__ void k_topk_reset(TopkState* st, int k) {
    st->prefix = 0; st->k_rem = k; st->count = 0; st->pad = 0;
}

// dense b0 maintenance: zero old support entries, write new ones
__global__ void k_scatter(double* __restrict__ dst, const int64_t* __restrict__ idx, const double* __restrict__ val,
                          int64_t k, int zero_only) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t < k) dst[idx[t]] = zero_only ? 0.0 : val[t];
}

static void topk_run(TopkCtx& c, const double* d_dfa, const double* d_b0d, const double* d_sinv, int64_t p_mod,
                     const double* d_bounds, double eta, double bound, int64_t k, cudaStream_t s,
                     const double* d_scal = nullptr, double bound_coef = 0.0, const double* d_l2 = nullptr) {
    int grid = tk_grid(c.p);
    int kk = (int)(k < c.p ? k : c.p);
    IHTB_LAUNCH(k_topk_reset, 1, 1, 0, s, c.st, kk);
    IHTB_LAUNCH(k_keys_hist0, grid, TK_THREADS, 0, s, c.p, d_dfa, d_b0d, d_sinv, p_mod, d_bounds, eta, bound, d_scal,
                bound_coef, c.wt, c.keyL, c.keyU, c.hist, d_l2);
    IHTB_LAUNCH(k_pick, 1, 1024, 0, s, c.hist, c.st, 21, 11);
    IHTB_LAUNCH(k_hist, grid, TK_THREADS, 0, s, c.p, c.keyL, c.st, 10, 11, c.hist);
    IHTB_LAUNCH(k_pick, 1, 1024, 0, s, c.hist, c.st, 10, 11);
    IHTB_LAUNCH(k_hist, grid, TK_THREADS, 0, s, c.p, c.keyL, c.st, 0, 10, c.hist);
    IHTB_LAUNCH(k_pick, 1, 1024, 0, s, c.hist, c.st, 0, 10);
    IHTB_LAUNCH(k_compact, grid, TK_THREADS, 0, s, c.p, c.keyU, c.st, c.cand, c.cap);
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
    IHTB_CUDA(cudaMemsetAsync(c.cand, 0xFF, (size_t)c.cap * sizeof(int64_t), s));
    topk_run(c, d_dfa, nullptr, d_sinv, c.p, nullptr, 1.0, (d_scal || d_l2) ? 0.0 : host_bound, k, s, d_scal, bound_coef,
             d_l2);
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
