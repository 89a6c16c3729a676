// Device side of the doubly sparse (group) projection: per-group candidate lists from the sweep's gradient.
//
// project_group_sparse!(b0 + eta*df) keeps, inside each of the J groups with the largest norm, the k_g largest
// entries.  Whatever eta is, the survivors of group g lie in supp(b0) plus the k_g largest |df| of g outside the
// support -- a subset of the 2 k_g largest |df| of g, because a group holds at most k_g support entries -- and a group
// without support entries has norm eta^2 * T_g with T_g = sum of its k_g largest df^2.  So once per sweep the device
// produces, for every group, that short list and T_g; the host picks the groups that can matter (those holding support
// entries and the J largest T_g among the others), the listed columns are re-scored exactly in FP64 (k_xt_gather),
// and the reference's projection then runs on the host on a few dozen exact values for the gradient step and every
// backtrack.  The FAST sweep's df carries an absolute error <= e_j = sinv_j * bound (sweep_lut.cu), so lists and
// group sums are taken with slack: with L_j = |df_j| - e_j, U_j = |df_j| + e_j and tau_g = the 2k_g-th largest L of
// the group, every member with U_j >= tau_g is listed (a superset of the true 2k_g largest), and
// T_L = sum of the k_g largest L^2 <= T_g <= sum of (L + 2 max_j e_j)^2 over the same entries = T_U.
#include "groups.cuh"

namespace ihtb {

constexpr int GT_THREADS = 128;
constexpr int GT_EXTRA = 8;          // list slots beyond 2 k_g for entries within the error bound of tau_g

__global__ void __launch_bounds__(GT_THREADS)
k_group_topk(const double* __restrict__ dfa, const double* __restrict__ sinv, const double* __restrict__ scal,
             double bound_coef, double bound, const int64_t* __restrict__ order, const int64_t* __restrict__ gptr,
             const int64_t* __restrict__ goff, const int64_t* __restrict__ ks, int64_t kscalar,
             const double* __restrict__ smax, int G, int64_t* __restrict__ out_idx, double* __restrict__ out_T) {
    __shared__ double s_a[GT_THREADS];
    __shared__ int64_t s_j[GT_THREADS];
    __shared__ int s_extra;
    if (scal) bound = bound_coef * (scal[1] + fabs(scal[0]));
    const int g = blockIdx.x;
    const int64_t m0 = gptr[g], m1 = gptr[g + 1];
    const int64_t kg = ks ? ks[g] : kscalar;
    const int64_t cap = goff[g + 1] - goff[g];
    int64_t take = 2 * kg < m1 - m0 ? 2 * kg : m1 - m0;
    if (take > cap) take = cap;
    const double emax = smax[g] * bound;
    auto lower = [&](int64_t j) -> double {
        const double a = fabs(dfa[j]);
        const double lo = a - (sinv[j] * bound + a * 4e-16);
        return lo > 0.0 ? lo : 0.0;                    // NaN -> 0
    };
    double prev_a = INFINITY;
    int64_t prev_j = -1;
    double TL = 0.0, TU = 0.0;
    int64_t t = 0;
    if (threadIdx.x == 0) s_extra = 0;
    __syncthreads();
    for (; t < take; ++t) {
        double best_a = -1.0;
        int64_t best_j = INT64_MAX;
        for (int64_t i = m0 + threadIdx.x; i < m1; i += GT_THREADS) {
            const int64_t j = order[i];
            const double a = lower(j);
            const bool eligible = (a < prev_a) || (a == prev_a && j > prev_j);
            if (eligible && (a > best_a || (a == best_a && j < best_j))) { best_a = a; best_j = j; }
        }
        s_a[threadIdx.x] = best_a; s_j[threadIdx.x] = best_j;
        __syncthreads();
        for (int o = GT_THREADS / 2; o >= 1; o >>= 1) {
            if (threadIdx.x < o) {
                const double a2 = s_a[threadIdx.x + o];
                const int64_t j2 = s_j[threadIdx.x + o];
                if (a2 > s_a[threadIdx.x] || (a2 == s_a[threadIdx.x] && j2 < s_j[threadIdx.x])) {
                    s_a[threadIdx.x] = a2; s_j[threadIdx.x] = j2;
                }
            }
            __syncthreads();
        }
        const double wa = s_a[0];
        const int64_t wj = s_j[0];
        __syncthreads();
        if (wa < 0.0) break;
        if (threadIdx.x == 0) {
            out_idx[goff[g] + t] = wj;
            if (t < kg) {
                const double up = wa + 2.0 * emax + (wa + 2.0 * emax) * 1e-15;
                TL += wa * wa;
                TU += up * up;
            }
        }
        prev_a = wa; prev_j = wj;
    }
    // entries after the cursor whose upper bound still reaches tau_g = prev_a (only if the group has more members)
    if (t == take && take < m1 - m0) {
        for (int64_t i = m0 + threadIdx.x; i < m1; i += GT_THREADS) {
            const int64_t j = order[i];
            const double a = fabs(dfa[j]);
            const double lo = lower(j);
            const bool after = (lo < prev_a) || (lo == prev_a && j > prev_j);
            double up = a + (sinv[j] * bound + a * 4e-16);
            if (!(up >= 0.0)) up = INFINITY;              // NaN: always a candidate
            if (after && up >= prev_a) {
                const int pos = atomicAdd(&s_extra, 1);
                if (take + pos < cap) out_idx[goff[g] + take + pos] = j;
                else out_T[2 * (int64_t)G] = 1.0;         // overflow: the host reports a degenerate projection
            }
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int64_t used = t + s_extra;
        if (used > cap) used = cap;
        for (int64_t u = used; u < cap; ++u) out_idx[goff[g] + u] = -1;
        out_T[g] = TL;
        out_T[G + g] = TU;
    }
}

__global__ void k_group_take(const int32_t* __restrict__ chosen, int n, int64_t lcap, const int64_t* __restrict__ goff,
                             const int64_t* __restrict__ gidx, int64_t* __restrict__ oidx) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= (int64_t)n * lcap) return;
    const int c = (int)(t / lcap);
    const int64_t u = t % lcap;
    const int g = chosen[c];
    const int64_t len = goff[g + 1] - goff[g];
    oidx[t] = (u < len) ? gidx[goff[g] + u] : -1;
}

// group1: 1-based group of every SNP of the WHOLE matrix (p_global entries); this handle owns columns [j0, j0+p).
// Group ids, check_group and the list capacity lcap are global; the member lists are those of the local columns.
void GroupCtx::build(int64_t p_, int64_t j0, int64_t p_global, const int32_t* group1, int J_, const int64_t* ks_,
                     int64_t n_groups, int64_t kscalar, const double* h_sinv) {
    p = p_; J = J_;
    int gmax = 0;
    for (int64_t j = 0; j < p_global; ++j) {
        IHTB_CHECK(group1[j] >= 1, IHTB_EDOMAIN, "group ids must be >= 1");
        gmax = std::max(gmax, (int)group1[j]);
    }
    G = gmax;
    ks_vector = ks_ != nullptr;
    if (ks_vector) {
        // check_group (src/utilities.jl:902-915)
        IHTB_CHECK(p_global > 1, IHTB_EINVAL, "Doubly sparse projection specified (since k is a vector) but there are no group information.");
        IHTB_CHECK(n_groups >= G, IHTB_EDIM, "k must have one entry per group");
        ks.assign(ks_, ks_ + n_groups);
        if ((int64_t)G < n_groups) G = (int)n_groups;
    } else {
        ks.clear();
    }
    kcap = kscalar;
    grp.resize((size_t)p_global);
    std::vector<int64_t> gsize_all((size_t)G, 0);
    gsize.assign((size_t)G, 0);
    for (int64_t j = 0; j < p_global; ++j) {
        grp[(size_t)j] = group1[j] - 1;
        ++gsize_all[(size_t)grp[(size_t)j]];
        if (j >= j0 && j < j0 + p) ++gsize[(size_t)grp[(size_t)j]];
    }
    if (ks_vector)
        for (int g = 0; g < G; ++g)
            IHTB_CHECK(gsize_all[(size_t)g] > ks[(size_t)g], IHTB_EDOMAIN,
                       "Maximum predictors for group " + std::to_string(g + 1) + " was " + std::to_string(ks[(size_t)g]) +
                           " but there are only " + std::to_string(gsize_all[(size_t)g]) +
                           " predictors is this group. Please choose a smaller number.");
    std::vector<int64_t> gptr((size_t)G + 1, 0), order((size_t)p);
    for (int g = 0; g < G; ++g) gptr[(size_t)g + 1] = gptr[(size_t)g] + gsize[(size_t)g];
    {
        std::vector<int64_t> fill(gptr.begin(), gptr.end() - 1);
        for (int64_t j = 0; j < p; ++j) order[(size_t)fill[(size_t)grp[(size_t)(j0 + j)]]++] = j;      // LOCAL column indices
    }
    goff.assign((size_t)G + 1, 0);
    lcap = 1;
    for (int g = 0; g < G; ++g) {
        const int64_t want = 2 * k_of(g, kscalar) + GT_EXTRA;
        goff[(size_t)g + 1] = goff[(size_t)g] + std::min<int64_t>(want, gsize[(size_t)g]);
        lcap = std::max(lcap, std::min<int64_t>(want, gsize_all[(size_t)g]));      // the same on every rank
    }
    const int64_t L = std::max<int64_t>(goff[(size_t)G], 1);
    d_order.alloc((size_t)std::max<int64_t>(p, 1)); d_gptr.alloc((size_t)G + 1); d_goff.alloc((size_t)G + 1);
    d_gidx.alloc((size_t)L); d_gT.alloc(2 * (size_t)G + 1); h_gT.alloc(2 * (size_t)G + 1);
    std::vector<double> smax((size_t)G, 0.0);
    for (int64_t j = 0; j < p; ++j) {
        double& m = smax[(size_t)grp[(size_t)(j0 + j)]];
        m = std::max(m, h_sinv[j]);
    }
    d_smax.alloc((size_t)G);
    IHTB_CUDA(cudaMemcpy(d_smax.p, smax.data(), (size_t)G * sizeof(double), cudaMemcpyHostToDevice));
    IHTB_CUDA(cudaMemcpy(d_order.p, order.data(), (size_t)p * sizeof(int64_t), cudaMemcpyHostToDevice));
    IHTB_CUDA(cudaMemcpy(d_gptr.p, gptr.data(), ((size_t)G + 1) * sizeof(int64_t), cudaMemcpyHostToDevice));
    IHTB_CUDA(cudaMemcpy(d_goff.p, goff.data(), ((size_t)G + 1) * sizeof(int64_t), cudaMemcpyHostToDevice));
    if (ks_vector) {
        d_ks.alloc((size_t)G);
        IHTB_CUDA(cudaMemcpy(d_ks.p, ks.data(), (size_t)G * sizeof(int64_t), cudaMemcpyHostToDevice));
    }
    chosen_cap = 0;
    ensure_chosen(std::min(G, 64));
}

void GroupCtx::ensure_chosen(int n) {
    if (n <= chosen_cap) return;
    chosen_cap = std::max(n, 2 * chosen_cap);
    d_chosen.alloc((size_t)chosen_cap);
    d_oidx.alloc((size_t)chosen_cap * lcap); d_oval.alloc((size_t)chosen_cap * lcap);
    h_oidx.alloc((size_t)chosen_cap * lcap); h_oval.alloc((size_t)chosen_cap * lcap);
}

void group_topk(GroupCtx& c, const double* d_dfa, const double* d_sinv, const double* d_scal, double bound_coef,
                double host_bound, int64_t kscalar, cudaStream_t s) {
    IHTB_CUDA(cudaMemsetAsync(c.d_gT.p + 2 * (size_t)c.G, 0, sizeof(double), s));
    IHTB_LAUNCH(k_group_topk, c.G, GT_THREADS, 0, s, d_dfa, d_sinv, d_scal, bound_coef, host_bound, c.d_order.p,
                c.d_gptr.p, c.d_goff.p, c.ks_vector ? c.d_ks.p : (const int64_t*)nullptr, kscalar, c.d_smax.p, c.G,
                c.d_gidx.p, c.d_gT.p);
}

// sharded exchange block: [global index or -1 | bits of the exact value] for `slots` gathered columns
__global__ void k_group_pack(const int64_t* __restrict__ oidx, const double* __restrict__ oval, int64_t slots,
                             int64_t j0, int64_t* __restrict__ block) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= slots) return;
    const int64_t j = oidx[t];
    block[t] = j >= 0 ? j + j0 : -1;
    block[slots + t] = j >= 0 ? __double_as_longlong(oval[t]) : 0;
}
void group_pack(const int64_t* d_oidx, const double* d_oval, int64_t slots, int64_t j0, int64_t* d_block, cudaStream_t s) {
    if (slots) IHTB_LAUNCH(k_group_pack, (unsigned)ceil_div(slots, 128), 128, 0, s, d_oidx, d_oval, slots, j0, d_block);
}

void group_take(GroupCtx& c, int n, cudaStream_t s) {
    if (n == 0) return;
    IHTB_LAUNCH(k_group_take, (unsigned)ceil_div((int64_t)n * c.lcap, 128), 128, 0, s, c.d_chosen.p, n, c.lcap,
                c.d_goff.p, c.d_gidx.p, c.d_oidx.p);
}

}  // namespace ihtb
