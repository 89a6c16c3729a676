// Device side of the doubly sparse (group) projection: per-group candidate lists from the exact gradient.
//
// project_group_sparse!(b0 + eta*df) keeps, inside each of the J groups with the largest norm, the k_g largest
// entries.  Whatever eta is, the survivors of group g lie in supp(b0) plus the k_g largest |df| of g outside the
// support -- a subset of the 2 k_g largest |df| of g, because a group holds at most k_g support entries -- and a group
// without support entries has norm eta^2 * T_g with T_g = sum of its k_g largest df^2.  So once per sweep the device
// produces, for every group, that short list and T_g; the host then picks the groups that can matter (those holding
// support entries and the J largest T_g among the others) and runs the reference's projection on a few dozen exact
// values for the gradient step and every backtrack.  Group fits use the exact FP64 sweep, so no error bounds are needed.
#include "groups.cuh"

namespace ihtb {

constexpr int GT_THREADS = 128;

__global__ void __launch_bounds__(GT_THREADS)
k_group_topk(const double* __restrict__ dfa, const int64_t* __restrict__ order, const int64_t* __restrict__ gptr,
             const int64_t* __restrict__ goff, const int64_t* __restrict__ ks, int64_t kscalar,
             int64_t* __restrict__ out_idx, double* __restrict__ out_val, double* __restrict__ out_T) {
    __shared__ double s_a[GT_THREADS];
    __shared__ int64_t s_j[GT_THREADS];
    const int g = blockIdx.x;
    const int64_t m0 = gptr[g], m1 = gptr[g + 1];
    const int64_t kg = ks ? ks[g] : kscalar;
    const int64_t cap = goff[g + 1] - goff[g];
    int64_t take = 2 * kg < m1 - m0 ? 2 * kg : m1 - m0;
    if (take > cap) take = cap;
    double prev_a = INFINITY;
    int64_t prev_j = -1;
    double T = 0.0;
    int64_t t = 0;
    for (; t < take; ++t) {
        double best_a = -1.0;
        int64_t best_j = INT64_MAX;
        for (int64_t i = m0 + threadIdx.x; i < m1; i += GT_THREADS) {
            const int64_t j = order[i];
            const double a = fabs(dfa[j]);
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
        if (wa < 0.0) break;                       // nothing eligible is left (NaN entries are never eligible)
        if (threadIdx.x == 0) {
            const double v = dfa[wj];
            out_idx[goff[g] + t] = wj;
            out_val[goff[g] + t] = v;
            if (t < kg) T += v * v;
        }
        prev_a = wa; prev_j = wj;
    }
    if (threadIdx.x == 0) {
        for (int64_t u = t; u < cap; ++u) { out_idx[goff[g] + u] = -1; out_val[goff[g] + u] = 0.0; }
        out_T[g] = T;
    }
}

__global__ void k_group_take(const int32_t* __restrict__ chosen, int n, int64_t lcap, const int64_t* __restrict__ goff,
                             const int64_t* __restrict__ gidx, const double* __restrict__ gval,
                             int64_t* __restrict__ oidx, double* __restrict__ oval) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= (int64_t)n * lcap) return;
    const int c = (int)(t / lcap);
    const int64_t u = t % lcap;
    const int g = chosen[c];
    const int64_t len = goff[g + 1] - goff[g];
    oidx[t] = (u < len) ? gidx[goff[g] + u] : -1;
    oval[t] = (u < len) ? gval[goff[g] + u] : 0.0;
}

void GroupCtx::build(int64_t p_, const int32_t* group1, int J_, const int64_t* ks_, int64_t n_groups, int64_t kscalar) {
    p = p_; J = J_;
    int gmax = 0;
    for (int64_t j = 0; j < p; ++j) {
        IHTB_CHECK(group1[j] >= 1, IHTB_EDOMAIN, "group ids must be >= 1");
        gmax = std::max(gmax, (int)group1[j]);
    }
    G = gmax;
    ks_vector = ks_ != nullptr;
    if (ks_vector) {
        // check_group (src/utilities.jl:902-915)
        IHTB_CHECK(p > 1, IHTB_EINVAL, "Doubly sparse projection specified (since k is a vector) but there are no group information.");
        IHTB_CHECK(n_groups >= G, IHTB_EDIM, "k must have one entry per group");
        ks.assign(ks_, ks_ + n_groups);
        if ((int64_t)G < n_groups) G = (int)n_groups;
    } else {
        ks.clear();
    }
    kcap = kscalar;
    grp.resize((size_t)p);
    gsize.assign((size_t)G, 0);
    for (int64_t j = 0; j < p; ++j) { grp[(size_t)j] = group1[j] - 1; ++gsize[(size_t)grp[(size_t)j]]; }
    if (ks_vector)
        for (int g = 0; g < G; ++g)
            IHTB_CHECK(gsize[(size_t)g] > ks[(size_t)g], IHTB_EDOMAIN,
                       "Maximum predictors for group " + std::to_string(g + 1) + " was " + std::to_string(ks[(size_t)g]) +
                           " but there are only " + std::to_string(gsize[(size_t)g]) +
                           " predictors is this group. Please choose a smaller number.");
    std::vector<int64_t> gptr((size_t)G + 1, 0), order((size_t)p);
    for (int g = 0; g < G; ++g) gptr[(size_t)g + 1] = gptr[(size_t)g] + gsize[(size_t)g];
    {
        std::vector<int64_t> fill(gptr.begin(), gptr.end() - 1);
        for (int64_t j = 0; j < p; ++j) order[(size_t)fill[(size_t)grp[(size_t)j]]++] = j;
    }
    goff.assign((size_t)G + 1, 0);
    lcap = 1;
    for (int g = 0; g < G; ++g) {
        const int64_t len = std::min<int64_t>(2 * k_of(g, kscalar), gsize[(size_t)g]);
        goff[(size_t)g + 1] = goff[(size_t)g] + len;
        lcap = std::max(lcap, len);
    }
    const int64_t L = std::max<int64_t>(goff[(size_t)G], 1);
    d_order.alloc((size_t)p); d_gptr.alloc((size_t)G + 1); d_goff.alloc((size_t)G + 1);
    d_gidx.alloc((size_t)L); d_gval.alloc((size_t)L); d_gT.alloc((size_t)G); h_gT.alloc((size_t)G);
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

void group_topk(GroupCtx& c, const double* d_dfa, int64_t kscalar, cudaStream_t s) {
    IHTB_LAUNCH(k_group_topk, c.G, GT_THREADS, 0, s, d_dfa, c.d_order.p, c.d_gptr.p, c.d_goff.p,
                c.ks_vector ? c.d_ks.p : (const int64_t*)nullptr, kscalar, c.d_gidx.p, c.d_gval.p, c.d_gT.p);
}

void group_take(GroupCtx& c, int n, cudaStream_t s) {
    if (n == 0) return;
    IHTB_LAUNCH(k_group_take, (unsigned)ceil_div((int64_t)n * c.lcap, 128), 128, 0, s, c.d_chosen.p, n, c.lcap,
                c.d_goff.p, c.d_gidx.p, c.d_gval.p, c.d_oidx.p, c.d_oval.p);
}

}  // namespace ihtb
