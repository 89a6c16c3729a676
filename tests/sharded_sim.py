"""numpy model of the SNP-sharded IHT protocol that fit.cu implements, run over torch.distributed
(gloo on CPU).  Each rank owns a column block; partial X*coef n-vectors are all-reduced, the global top-k is taken
from the all-gathered local top-k candidates, and values of other shards' columns are filled in by a zero-padded
all-reduce.  Used by tests/test_multiproc_cpu.py to show the protocol reproduces the single-process oracle."""
import numpy as np
import torch
import torch.distributed as dist

from oracle import glm
from oracle.iht import IHTTrace, fit_iht_loop, project_group_sparse


def _allreduce(a):
    t = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64))
    dist.all_reduce(t)
    return t.numpy()


class ShardedIHT:
    """Same state machine as oracle.iht.IHTVariable (memory-efficient branch), sharded by columns."""

    def __init__(self, x_local, j0, p_global, z, y, k, d, l, J=1, group=None):
        self.x, self.j0, self.p, self.pl = x_local, j0, p_global, x_local.shape[1]
        self.J, self.group = J, (None if group is None else np.asarray(group, dtype=np.int64))   # 1-based ids, p_global
        self.y = np.asarray(y, float); self.z = z.reshape(-1, 1) if z.ndim == 1 else z
        self.n, self.q, self.k, self.d, self.l = self.y.shape[0], self.z.shape[1], k, d, l
        self.est_r, self.nb_r = "None", 1.0
        self.world = dist.get_world_size()

    def local(self, j):
        return (j >= self.j0) & (j < self.j0 + self.pl)

    def support_matvec(self, idx, coef):
        m = self.local(idx) & (coef != 0)
        part = self.x.support_xb(idx[m] - self.j0, coef[m]) if m.any() else np.zeros(self.n)
        return _allreduce(part)

    def gather_values(self, cols):
        """df at global columns `cols` (zero-padded all-reduce: exact, other shards add 0)."""
        out = np.zeros(len(cols))
        m = self.local(cols)
        out[m] = self.df_local[cols[m] - self.j0]
        return _allreduce(out)

    def global_topk(self, v_local, extra_idx):
        """Indices of the global top-k of |v| given each rank's local vector: union of local top-k's (+ extra)."""
        kk = min(self.k, self.pl)
        loc = np.argsort(-np.abs(v_local), kind="stable")[:kk] + self.j0
        buf = np.full(self.k, -1, dtype=np.int64); buf[:kk] = loc
        allb = [torch.zeros(self.k, dtype=torch.int64) for _ in range(self.world)]
        dist.all_gather(allb, torch.from_numpy(buf))
        cand = np.concatenate([b.numpy() for b in allb] + [np.asarray(extra_idx, dtype=np.int64)])
        return np.unique(cand[cand >= 0])

    def group_candidates(self):
        """Doubly sparse fits (fit.cu select_groups): every rank lists, per group, its 2k largest local |df| and the
        sum T of its k largest df^2; the bounds are gathered and combined (lower = max over ranks, upper = sum over
        ranks), the groups that can reach the J best norms are chosen identically on every rank, and their local lists
        are all-gathered."""
        G = int(self.group.max())
        gl = self.group[self.j0:self.j0 + self.pl]
        lists, T = [], np.zeros(G)
        for g in range(1, G + 1):
            mem = np.flatnonzero(gl == g)
            order = mem[np.lexsort((mem, -np.abs(self.df_local[mem])))]
            lists.append(order[:2 * self.k] + self.j0)
            T[g - 1] = float(np.sum(self.df_local[order[:self.k]] ** 2))
        allT = [torch.zeros(G, dtype=torch.float64) for _ in range(self.world)]
        dist.all_gather(allT, torch.from_numpy(T))
        TL = np.max([t.numpy() for t in allT], axis=0); TU = np.sum([t.numpy() for t in allT], axis=0)
        has = np.zeros(G, bool); has[self.group[np.flatnonzero(self.b0)] - 1] = True
        chosen = list(np.flatnonzero(has))
        lows = TL[~has]
        if lows.size and self.J > 0:
            thr = np.sort(lows)[-min(self.J, lows.size)]
            chosen += [g for g in np.flatnonzero(~has) if TU[g] >= thr]
        width = 2 * self.k * max(len(chosen), 1)
        buf = np.full(width, -1, dtype=np.int64)
        mine = np.concatenate([lists[g] for g in chosen]) if chosen else np.zeros(0, dtype=np.int64)
        buf[:mine.size] = mine
        allb = [torch.zeros(width, dtype=torch.int64) for _ in range(self.world)]
        dist.all_gather(allb, torch.from_numpy(buf))
        cand = np.concatenate([b.numpy() for b in allb])
        return np.unique(cand[cand >= 0])

    # ---- the oracle's interface ---------------------------------------------------------------
    def init_iht_indices(self, cv_idx):
        self.b = np.zeros(self.p); self.b0 = np.zeros(self.p); self.best_b = np.zeros(self.p)
        self.c = np.zeros(self.q); self.c0 = np.zeros(self.q); self.best_c = np.zeros(self.q)
        self.idx = np.zeros(self.p, bool); self.idx0 = np.zeros(self.p, bool)
        self.idc = np.ones(self.q, bool)
        self.xb = np.zeros(self.n)
        self.cv_wts = np.asarray(cv_idx, float)
        ybar = float(np.sum(self.y * self.cv_wts)) / int(np.count_nonzero(self.cv_wts))
        for _ in range(20):
            g1 = float(glm.linkinv(self.l, self.c[0])); g2 = float(glm.mueta(self.l, self.c[0]))
            self.c[0] -= min(max((g1 - ybar) / g2, -1.0), 1.0)
            if abs(g1 - ybar) < 1e-10:
                break
        self.zc = self.z @ self.c
        self.update_mu(); self.score()
        cand = self.global_topk(self.df_local, [])
        vals = self.gather_values(cand)
        order = sorted(range(len(cand)), key=lambda t: (-abs(vals[t]), cand[t]))[: self.k]
        self.df_sparse = {int(cand[t]): float(vals[t]) for t in order if vals[t] != 0}
        self.idx[:] = False; self.idx[list(self.df_sparse)] = True
        self.use_sparse_df = True

    def df_at(self, cols):
        if self.use_sparse_df:
            return np.array([self.df_sparse.get(int(j), 0.0) for j in cols])
        return self.gather_values(np.asarray(cols, dtype=np.int64))

    def update_mu(self):
        self.mu = glm.linkinv(self.l, self.xb + self.zc)

    def update_xb(self):
        idx = np.flatnonzero(self.idx)
        self.xb = self.support_matvec(idx, self.b[idx])
        self.zc = self.z @ self.c
        if self.d != glm.NORMAL:
            np.clip(self.xb, -20, 20, out=self.xb); np.clip(self.zc, -20, 20, out=self.zc)

    def loglikelihood(self):
        return glm.loglikelihood(self.d, self.y, self.mu, self.cv_wts, self.nb_r)

    def score(self):
        with np.errstate(divide="ignore", invalid="ignore"):
            w = glm.mueta(self.l, self.xb + self.zc) / glm.glmvar(self.d, self.mu, self.nb_r)
            self.r = w * (self.y - self.mu) * self.cv_wts
        self.df_local = self.x.xt_v(self.r)
        self.df2 = self.z.T @ self.r
        self.use_sparse_df = False

    def iht_stepsize(self):
        idx = np.flatnonzero(self.idx)
        dfi = self.df_at(idx)
        xgk = self.support_matvec(idx, dfi) + self.z[:, self.idc] @ self.df2[self.idc]
        with np.errstate(divide="ignore", invalid="ignore"):
            sw = np.sqrt(glm.mueta(self.l, self.xb + self.zc) ** 2 / glm.glmvar(self.d, self.mu, self.nb_r)) * self.cv_wts
            xgk = xgk * sw
            eta = (float(np.sum(dfi ** 2)) + float(np.sum(self.df2[self.idc] ** 2))) / float(np.dot(xgk, xgk))
        return 1e-8 if (np.isinf(eta) or np.isnan(eta)) else eta

    def iht_gradstep(self, eta):
        supp0 = np.flatnonzero(self.b0)
        if self.group is not None:
            if self.use_sparse_df:
                cand = np.unique(np.concatenate([supp0, np.fromiter(self.df_sparse, dtype=np.int64)]))
            else:
                cand = np.unique(np.concatenate([supp0, self.group_candidates()])).astype(np.int64)
            self.b = np.zeros(self.p)
            self.b[cand] = self.b0[cand] + eta * self.df_at(cand)
            project_group_sparse(self.b, self.group, self.J, self.k)       # entries outside `cand` cannot survive
            self.c = self.c0 + eta * self.df2
            self.idx = self.b != 0; self.idc = self.c != 0
            return
        if self.use_sparse_df:
            cand = np.unique(np.concatenate([supp0, np.fromiter(self.df_sparse, dtype=np.int64)]))
        else:
            v_local = self.b0[self.j0:self.j0 + self.pl] + eta * self.df_local
            cand = self.global_topk(v_local, supp0)
        v = self.b0[cand] + eta * self.df_at(cand)
        order = sorted(range(len(cand)), key=lambda t: (-abs(v[t]), cand[t]))[: self.k]
        self.b = np.zeros(self.p)
        for t in order:
            self.b[cand[t]] = v[t]
        self.c = self.c0 + eta * self.df2
        self.idx = self.b != 0; self.idc = self.c != 0

    def save_prev(self, cur, best):
        self.b0 = self.b.copy(); self.c0 = self.c.copy(); self.idx0 = self.idx.copy()
        if cur > best:
            self.best_b = self.b.copy(); self.best_c = self.c.copy()
        return max(cur, best)

    def check_convergence(self):
        nrm = max(np.max(np.abs(self.b - self.b0)), np.max(np.abs(self.c - self.c0)))
        return nrm / (max(np.max(np.abs(self.b0)), np.max(np.abs(self.c0))) + 1.0)

    def backtrack(self, eta):
        self.b = self.b0.copy(); self.c = self.c0.copy()
        self.iht_gradstep(eta); self.update_xb(); self.update_mu()
        return self.loglikelihood()

    def save_best_model(self):
        self.b = self.best_b.copy(); self.c = self.best_c.copy()
        self.idx = self.b != 0; self.idc = self.c != 0
        self.update_xb()
        self.mu = glm.linkinv(self.l, self.xb)


def fit_sharded(y, x_local, j0, p_global, z, k, d, l, max_iter=200, J=1, group=None):
    v = ShardedIHT(x_local, j0, p_global, z, y, k, d, l, J, group)
    v.init_iht_indices(np.ones(v.n, bool))
    tr = IHTTrace()
    best, it = fit_iht_loop(v, max_iter=max_iter, trace=tr)
    return v, best, it, tr
