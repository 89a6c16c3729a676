"""CPU property tests of the arguments the device-side candidate selection rests on (DESIGN.md sections 3.2, 6, 7).
They restate the selection rules of topk.cu / groups.cu / fit.cu in numpy on random inputs and check, against the
oracle's projections, that the short candidate lists always contain what the full projection keeps."""
import numpy as np
import pytest

from oracle import iht


def _topk_support(v, k):
    x = v.copy()
    iht.project_k(x, k)
    return set(np.flatnonzero(x))


@pytest.mark.parametrize("seed", range(20))
def test_eta_independent_candidates_contain_every_projection(seed):
    """supp(P_k(b0 + eta*df)) lies in supp(b0) + the k largest |df| outside supp(b0), whatever eta is
    (one selection per sweep serves the gradient step and all backtracks; fit.cu select_rescore)."""
    rng = np.random.default_rng(seed)
    p, k = 400, int(rng.integers(1, 12))
    b0 = np.zeros(p)
    supp = rng.choice(p, size=int(rng.integers(0, k + 1)), replace=False)
    b0[supp] = rng.normal(size=supp.size)
    df = rng.normal(size=p) * rng.choice([1e-3, 1.0, 50.0])
    outside = np.setdiff1d(np.arange(p), supp)
    cand = set(supp) | set(outside[np.argsort(-np.abs(df[outside]), kind="stable")[:k]])
    for eta in (1e-8, 1e-3, 0.37, 1.0, 20.0, 1e4):
        assert _topk_support(b0 + eta * df, k) <= cand


@pytest.mark.parametrize("seed", range(20))
def test_sweep_error_bound_selection_contains_true_topk(seed):
    """topk.cu: with |approx_j - true_j| <= e_j, tau = k-th largest (|approx| - e) and the candidates {|approx| + e >= tau}
    contain the true top-k."""
    rng = np.random.default_rng(100 + seed)
    p, k = 1000, int(rng.integers(1, 30))
    true = rng.normal(size=p)
    e = np.abs(rng.normal(size=p)) * rng.choice([1e-6, 1e-2, 0.3])
    approx = true + e * rng.uniform(-1, 1, size=p)
    lo = np.maximum(np.abs(approx) - e, 0.0); up = np.abs(approx) + e
    tau = np.sort(lo)[-k]
    cand = set(np.flatnonzero(up >= tau))
    assert set(np.argsort(-np.abs(true), kind="stable")[:k]) <= cand


def _group_lists(approx, e, members, kg, extra=8):
    """groups.cu k_group_topk for one group on one rank: (list, T_L, T_U) from approximate values with error e."""
    if members.size == 0:
        return [], 0.0, 0.0
    a = np.abs(approx[members]); err = e[members]
    lo = np.maximum(a - err, 0.0); up = a + err
    order = np.lexsort((members, -lo))                  # (L desc, index asc)
    take = min(2 * kg, members.size)
    first = order[:take]
    lst = list(members[first])
    emax = err.max()
    TL = float(np.sum(lo[first[:kg]] ** 2)); TU = float(np.sum((lo[first[:kg]] + 2 * emax) ** 2))
    if take < members.size:
        tau = lo[first[-1]]
        rest = order[take:]
        lst += list(members[rest[up[rest] >= tau]])
    return lst, TL, TU


@pytest.mark.parametrize("seed", range(25))
def test_group_candidates_are_safe_also_when_sharded(seed):
    """The per-group lists and the T_L / T_U bounds of groups.cu, combined over R shards as max / sum (fit.cu
    select_groups), always retain every entry that project_group_sparse!(b0 + eta*df) keeps."""
    rng = np.random.default_rng(500 + seed)
    p, G, J, k = 600, int(rng.integers(2, 12)), int(rng.integers(1, 4)), int(rng.integers(1, 4))
    R = int(rng.integers(1, 4))
    group = rng.integers(1, G + 1, size=p)
    group[:G] = np.arange(1, G + 1)                     # every group has a member
    true = rng.normal(size=p) * rng.choice([1e-2, 1.0])
    e = np.abs(rng.normal(size=p)) * rng.choice([1e-7, 1e-3, 0.05])
    approx = true + e * rng.uniform(-1, 1, size=p)
    # previous iterate: a feasible point of the group projection
    b0 = rng.normal(size=p)
    iht.project_group_sparse(b0, group, J, k)
    supp = np.flatnonzero(b0)
    bounds = np.linspace(0, p, R + 1).astype(int)
    TL = np.zeros(G); TU = np.zeros(G); lists = [[] for _ in range(G)]
    for r in range(R):
        loc = np.arange(bounds[r], bounds[r + 1])
        for g in range(G):
            lst, tl, tu = _group_lists(approx, e, loc[group[loc] == g + 1], k)
            lists[g] += lst; TL[g] = max(TL[g], tl); TU[g] += tu
    has = np.zeros(G, bool); has[group[supp] - 1] = True
    chosen = set(np.flatnonzero(has))
    lows = TL[~has]
    if lows.size:
        thr = np.sort(lows)[-min(J, lows.size)]
        chosen |= set(g for g in np.flatnonzero(~has) if TU[g] >= thr)
    cand = set(supp)
    for g in chosen:
        cand |= set(lists[g])
    for eta in (1e-6, 0.05, 1.0, 30.0):
        full = b0 + eta * true
        iht.project_group_sparse(full, group, J, k)
        assert set(np.flatnonzero(full)) <= cand
