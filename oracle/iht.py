"""Oracle (test infrastructure): univariate IHT, numpy restatement of the reference.

Follows, function by function (all paths under /root/reference):
  `fit_iht`            src/fit.jl:60-118
  `fit_iht!`           src/fit.jl:145-207
  `iht_one_step!`      src/fit.jl:213-263
  `init_iht_indices!`  src/utilities.jl:366-438
  `iht_stepsize!`      src/utilities.jl:722-764
  `_iht_gradstep!`     src/utilities.jl:252-280  (+ vectorize!/unvectorize! :291-354,
                        project_k! :553-559, _choose! :444-458)
  `update_xb!`         src/utilities.jl:93-118
  `update_mu!`         src/utilities.jl:74-82
  `loglikelihood` / `deviance`  src/utilities.jl:9-61
  `score!`             src/utilities.jl:126-135
  `save_prev!`         src/utilities.jl:702-712
  `check_convergence`  src/utilities.jl:953-957
  `backtrack!`         src/utilities.jl:959-973
  `save_best_model!`   src/utilities.jl:995-1006
  `mle_for_r`          src/utilities.jl:141-247
  `pve`                src/pve.jl:22-37

Documented deviation (shared with the CUDA path): when the k-th magnitude is tied, the reference
keeps every tie and then drops random entries of the support with the global RNG
(`_choose!`, src/utilities.jl:444-458).  Both the oracle and the product instead drop the tied
entries with the HIGHEST index, which is deterministic and identical whenever there is no tie.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np
from scipy.special import digamma, polygamma

from . import glm


def project_k(x: np.ndarray, k: int) -> None:
    """`project_k!(x, k)` (src/utilities.jl:553-559): zero everything with |x| < k-th largest |x|."""
    if k < 0:
        raise ValueError(f"Attempted to project to sparsity level {k}")
    ax = np.abs(x)
    a = np.partition(ax, ax.size - k)[ax.size - k]
    x[ax < a] = 0.0


def project_group_sparse(y: np.ndarray, group: np.ndarray, J: int, k) -> None:
    """`project_group_sparse!(y, group, J, k)` (src/utilities.jl:613-679), k an int or one int per group (1-based
    group ids).  Both sorts are stable, so equal magnitudes keep index order and equal group norms keep group order."""
    group = np.asarray(group, dtype=np.int64)
    groups = int(group.max())
    kk = np.full(groups, int(k), dtype=np.int64) if np.isscalar(k) else np.asarray(k, dtype=np.int64)
    perm = np.argsort(-np.abs(y), kind="stable")
    group_count = np.zeros(groups, dtype=np.int64)
    group_norm = np.zeros(groups)
    for j in perm:
        g = group[j] - 1
        if group_count[g] < kk[g]:
            group_norm[g] += y[j] ** 2
            group_count[g] += 1
    order = np.argsort(-group_norm, kind="stable")
    group_rank = np.empty(groups, dtype=np.int64)
    group_rank[order] = np.arange(1, groups + 1)
    group_count[:] = 1
    for j in perm:
        g = group[j] - 1
        if group_rank[g] > J or group_count[g] > kk[g]:
            y[j] = 0.0
        else:
            group_count[g] += 1


def check_group(k, group) -> None:
    """`check_group` (src/utilities.jl:902-915)."""
    if not np.isscalar(k):
        if group is None or len(group) <= 1:
            raise AssertionError("Doubly sparse projection specified (since k is a vector) but there are no group information.")
        group = np.asarray(group)
        for i, ki in enumerate(k, start=1):
            members = int((group == i).sum())
            if not members > ki:
                raise ValueError(f"Maximum predictors for group {i} was {ki} but there are only {members} predictors is "
                                 "this group. Please choose a smaller number.")
    elif k < 0:
        raise AssertionError("Value of k (max predictors per group) must be nonnegative!")


def linreg_columns(xs: np.ndarray, ys: np.ndarray):
    """`linreg!` (src/utilities.jl:824-842) for every column of xs [N, m] at once: regression of ys on [1, x] through
    the 2x2 Cholesky factor.  A failed factorisation (constant column) leaves xty unsolved: slope = sum(x y),
    intercept = sum(y) (the reference's try/catch, :836-841).  Returns (intercepts [m], slopes [m])."""
    N = float(xs.shape[0])
    sy = float(ys.sum())
    sx = xs.sum(axis=0); sxx = (xs * xs).sum(axis=0); sxy = xs.T @ ys
    icpt = np.full(xs.shape[1], sy); slope = sxy.copy()
    with np.errstate(divide="ignore", invalid="ignore"):
        u11 = np.sqrt(N); u12 = sx / u11
        d = sxx - u12 * u12
        ok = d > 0
        u22 = np.sqrt(np.where(ok, d, 1.0))
        t1 = sy / u11
        t2 = (sxy - u12 * t1) / u22
        b2 = t2 / u22
        b1 = (t1 - u12 * b2) / u11
    icpt[ok] = b1[ok]; slope[ok] = b2[ok]
    return icpt, slope


def prune_ties(vals: np.ndarray, nz_mask: np.ndarray, excess: int) -> None:
    """Deterministic replacement for `_choose!`: drop `excess` smallest-magnitude support entries,
    highest index first (only ties at the threshold can be in excess)."""
    pos = np.flatnonzero(nz_mask)
    order = sorted(pos, key=lambda j: (abs(vals[j]), -j))
    for j in order[:excess]:
        vals[j] = 0.0
        nz_mask[j] = False


@dataclass
class IHTTrace:
    logl: list = field(default_factory=list)
    backtracks: list = field(default_factory=list)
    tol: list = field(default_factory=list)
    eta: list = field(default_factory=list)
    support: list = field(default_factory=list)


@dataclass
class IHTResult:
    """`IHTResult` (src/data_structures.jl:245-258)."""
    time: float
    logl: float
    iter: int
    beta: np.ndarray
    c: np.ndarray
    J: int
    k: int
    group: list
    d: str
    sigma_g: float
    trace: IHTTrace = None
    r: float = 1.0   # NegativeBinomial nuisance parameter at exit


class IHTVariable:
    """`IHTVariable` (src/data_structures.jl:4-43), memory_efficient=true branch."""

    def __init__(self, x, z, y, k, d, l, zkeep=None, est_r="None", nb_r=1.0, weight=None, J=1, group=None):
        self.x, self.y = x, np.asarray(y, dtype=np.float64)
        z = np.asarray(z, dtype=np.float64)
        self.z = z.reshape(-1, 1) if z.ndim == 1 else z
        n, p = x.shape
        q = self.z.shape[1]
        if not (self.y.shape[0] == n == self.z.shape[0]):
            raise ValueError(f"row dimension of y, x, and z ({self.y.shape[0]}, {n}, {self.z.shape[0]}) are not equal")
        self.n, self.p, self.q = n, p, q
        # src/data_structures.jl:75-81: a vector k means per-group sparsity (`ks`), and then v.k = 0
        if np.isscalar(k):
            self.k, self.ks = int(k), None
        else:
            self.k, self.ks = 0, np.asarray(k, dtype=np.int64)
        self.J = int(J)
        self.group = None if group is None or len(group) == 0 else np.asarray(group, dtype=np.int64)
        if self.group is not None and self.group.shape[0] != p:
            raise ValueError(f"group must have length {p} but was {self.group.shape[0]}")
        self.d, self.l, self.est_r, self.nb_r = d, l, est_r, float(nb_r)
        self.zkeep = np.ones(q, dtype=bool) if zkeep is None else np.asarray(zkeep, dtype=bool)
        if self.zkeep.shape[0] != q:
            raise ValueError(f"zkeep must have length {q} but was {self.zkeep.shape[0]}")
        self.zkeepn = int(self.zkeep.sum())
        # prior weights scale b before the projection (src/data_structures.jl:36,71-73: length p).  The reference's
        # vectorize!/unvectorize! index weight[p+1:p+q] for the covariates, which is out of bounds under @inbounds
        # (src/utilities.jl:305-308,346-351); kept covariates overwrite that slot with Inf, so only projected covariates
        # (zkeep false) would see it.  Oracle and product define the covariate weights as 1.
        self.weight = None if weight is None or len(weight) == 0 else np.asarray(weight, dtype=np.float64)
        if self.weight is not None and self.weight.shape[0] != p:
            raise ValueError(f"weight must have length {p} but was {self.weight.shape[0]}")
        self.b = np.zeros(p); self.b0 = np.zeros(p); self.best_b = np.zeros(p)
        self.xb = np.zeros(n); self.xgk = np.zeros(n)
        self.idx = np.zeros(p, bool); self.idx0 = np.zeros(p, bool)
        self.idc = self.zkeep.copy(); self.idc0 = self.zkeep.copy()
        self.r = np.zeros(n); self.df = np.zeros(p); self.df2 = np.zeros(q)
        self.c = np.zeros(q); self.c0 = np.zeros(q); self.best_c = np.zeros(q)
        self.zc = np.zeros(n); self.mu = np.zeros(n); self.cv_wts = np.zeros(n)

    # -- src/utilities.jl:74-82
    def update_mu(self):
        self.mu = glm.linkinv(self.l, self.xb + self.zc)

    # -- src/utilities.jl:93-118
    def update_xb(self):
        idx = np.flatnonzero(self.idx)
        self.xb = self.x.support_xb(idx, self.b[idx])
        self.zc = self.z @ self.c
        if self.d != glm.NORMAL:
            np.clip(self.xb, -20, 20, out=self.xb)
            np.clip(self.zc, -20, 20, out=self.zc)

    # -- src/utilities.jl:52-61, 9-20
    def deviance(self):
        return glm.deviance(self.d, self.y, self.mu, self.cv_wts, self.nb_r)

    def loglikelihood(self):
        return glm.loglikelihood(self.d, self.y, self.mu, self.cv_wts, self.nb_r)

    # -- src/utilities.jl:126-135
    def score(self):
        eta = self.xb + self.zc
        with np.errstate(divide="ignore", invalid="ignore"):
            w = glm.mueta(self.l, eta) / glm.glmvar(self.d, self.mu, self.nb_r)
            self.r = w * (self.y - self.mu) * self.cv_wts
        self.df = self.x.xt_v(self.r)
        self.df2 = self.z.T @ self.r

    # -- src/utilities.jl:291-354 + 553-559 + 444-458 (no groups)
    def _project_full(self, b, c):
        """vectorize! -> project_k!(k + zkeepn) -> unvectorize!; returns nothing (in place)."""
        bw = b if self.weight is None else b * self.weight
        full = np.concatenate([bw, np.where(self.zkeep, np.inf, c)])
        project_k(full, self.k + self.zkeepn)
        b[:] = full[: self.p] if self.weight is None else full[: self.p] / self.weight
        cpart = full[self.p:]
        c[~self.zkeep] = cpart[~self.zkeep]

    def _choose(self):
        sparsity = self.k + self.zkeepn
        groups = 1 if self.J == 0 else self.J
        nonzero = int(self.idx.sum()) + int(self.idc.sum()) - self.zkeepn
        if nonzero > groups * sparsity:
            prune_ties(self.b, self.idx, nonzero - groups * sparsity)

    # -- src/utilities.jl:366-438
    # -- src/utilities.jl:776-842 (`initialize_beta!` + `linreg!`)
    def initialize_beta(self, cv_idx):
        """Univariate regression of y on [1, x_i] over the training samples for every SNP and covariate.
        A failed 2x2 Cholesky (monomorphic column in the training set) leaves xty unsolved: beta_i = sum(x y),
        intercept contribution = sum(y) (the reference's try/catch, :836-841)."""
        cv = np.asarray(cv_idx, dtype=bool)
        ys = self.y[cv]

        def linreg(xs):
            return linreg_columns(xs, ys)

        c0 = 0.0
        for j0 in range(0, self.p, 2048):     # column blocks bound the memory of the dense slice
            xs = self.x.columns(np.arange(j0, min(j0 + 2048, self.p)))[cv]
            icpt, slope = linreg(xs)
            c0 += float(icpt.sum())
            self.b[j0:j0 + 2048] = slope
        if self.q > 1:
            icpt, slope = linreg(self.z[cv][:, 1:])
            c0 += float(icpt.sum())
            self.c[1:] = slope
        self.c[0] = c0 / (self.p + self.q - 1)
        np.clip(self.b, -2, 2, out=self.b)
        np.clip(self.c, -2, 2, out=self.c)
        self.b0 = self.b.copy(); self.c0 = self.c.copy()

    def init_iht_indices(self, cv_idx: np.ndarray, init_beta: bool = False):
        for a in (self.b, self.b0, self.best_b, self.xb, self.xgk, self.r, self.df, self.df2, self.c,
                  self.best_c, self.c0, self.zc, self.mu, self.cv_wts):
            a[...] = 0
        self.idx[:] = False; self.idx0[:] = False
        self.idc = self.zkeep.copy(); self.idc0 = self.zkeep.copy()
        self.cv_wts[np.asarray(cv_idx, dtype=bool)] = 1.0
        # intercept by (clamped) Newton, :394-405
        ybar = float(np.sum(self.y * self.cv_wts)) / int(np.count_nonzero(self.cv_wts))
        for _ in range(20):
            g1 = float(glm.linkinv(self.l, self.c[0]))
            g2 = float(glm.mueta(self.l, self.c[0]))
            self.c[0] = self.c[0] - min(max((g1 - ybar) / g2, -1.0), 1.0)
            if abs(g1 - ybar) < 1e-10:
                break
        self.zc = self.z @ self.c
        self.update_mu()
        self.score()
        if init_beta:                          # :412-414 (df keeps the full gradient, xb / zc / mu are not refreshed)
            if self.d != glm.NORMAL:
                raise ValueError("Intializing beta values only work for Gaussian phenotypes! Sorry!")
            self.initialize_beta(cv_idx)
            self._project_full(self.b, self.c)            # project_k!(v) :561-573
            self.idx = self.b != 0
            self.idc = self.c != 0
            return
        if self.ks is not None:
            # :426-430: per-group sparsity projects df by groups, then reads the support off v.b -- which is all zero
            # at this point, so the fit starts from an EMPTY support with the group-projected gradient.
            project_group_sparse(self.df, self.group, self.J, self.ks)
            self.idx = self.b != 0
            self.idc = np.ones(self.q, dtype=bool)
            return
        # first k non-zero entries chosen from largest gradient; df itself is projected (:417-425)
        self._project_full(self.df, self.df2)
        self.idx = self.df != 0
        self.idc = self.zkeep.copy()
        sparsity = (self.k + self.zkeepn) * (1 if self.J == 0 else self.J)
        nonzero = int(self.idx.sum()) + int(self.idc.sum()) - self.zkeepn
        if nonzero > sparsity:
            # `_choose!` zeroes v.b (already 0) and clears idx; df keeps its value (:450-456)
            tmp = self.df.copy()
            prune_ties(tmp, self.idx, nonzero - sparsity)

    # -- src/utilities.jl:722-764
    def iht_stepsize(self):
        idx = np.flatnonzero(self.idx)
        self.xgk = self.x.support_xb(idx, self.df[idx])
        zdf2 = self.z[:, self.idc] @ self.df2[self.idc]
        self.xgk = self.xgk + zdf2
        with np.errstate(divide="ignore", invalid="ignore"):
            sw = np.sqrt(glm.mueta(self.l, self.xb + self.zc) ** 2 / glm.glmvar(self.d, self.mu, self.nb_r)) * self.cv_wts
            self.xgk = self.xgk * sw
            numer = float(np.sum(self.df[idx] ** 2)) + float(np.sum(self.df2[self.idc] ** 2))
            denom = float(np.dot(self.xgk, self.xgk))
            eta = numer / denom if denom != 0 else (np.inf if numer > 0 else np.nan)
        if np.isinf(eta) or np.isnan(eta):
            eta = 1e-8
        return eta

    # -- src/utilities.jl:252-280
    def iht_gradstep(self, eta):
        self.b += eta * self.df
        self.c += eta * self.df2
        if self.group is None:
            self._project_full(self.b, self.c)
        else:      # :267-269: groups project b only; no covariate selection, no weights
            project_group_sparse(self.b, self.group, self.J, self.k if self.ks is None else self.ks)
        self.idx = self.b != 0
        self.idc = self.c != 0
        if self.ks is None:
            self._choose()

    # -- src/utilities.jl:702-712
    def save_prev(self, cur_logl, best_logl):
        self.b0 = self.b.copy(); self.idx0 = self.idx.copy()
        self.idc0 = self.idc.copy(); self.c0 = self.c.copy()
        if cur_logl > best_logl:
            self.best_b = self.b.copy(); self.best_c = self.c.copy()
        return max(cur_logl, best_logl)

    # -- src/utilities.jl:953-957
    def check_convergence(self):
        the_norm = max(np.max(np.abs(self.b - self.b0)), np.max(np.abs(self.c - self.c0)))
        return the_norm / (max(np.max(np.abs(self.b0)), np.max(np.abs(self.c0))) + 1.0)

    # -- src/utilities.jl:959-973
    def backtrack(self, eta):
        self.b = self.b0.copy(); self.c = self.c0.copy()
        self.iht_gradstep(eta)
        self.update_xb(); self.update_mu()
        if self.est_r != "None":
            self.mle_for_r()
        return self.loglikelihood()

    # -- src/utilities.jl:995-1006
    def save_best_model(self):
        self.b = self.best_b.copy(); self.c = self.best_c.copy()
        self.idx = self.b != 0; self.idc = self.c != 0
        self.update_xb()
        self.mu = glm.linkinv(self.l, self.xb)   # genotype predictors only

    # -- src/utilities.jl:1014-1020 (memory_efficient=false branch: xk = x[:, idx], ALL n samples, no covariates)
    def debias(self):
        idx = np.flatnonzero(self.idx)
        if idx.size == 0:
            return
        xk = self.x.columns(idx)
        self.b[idx] = glm.glm_fit(xk, self.y, self.d, self.l, self.nb_r)

    # -- src/utilities.jl:141-247
    def mle_for_r(self):
        if self.est_r == "MM":
            self.nb_r = self._update_r_mm()
        elif self.est_r == "Newton":
            self.nb_r = self._update_r_newton()
        else:
            raise ValueError(f"Only support method is Newton or MM, but got {self.est_r}")

    def _update_r_mm(self):
        r = self.nb_r
        num = 0.0
        for yi in self.y:
            j = np.arange(0, int(yi))
            num += float(np.sum(r / (r + j)))
        den = float(np.sum(np.log(r / (r + self.mu))))
        return -num / den

    def _update_r_newton(self, max_iter=100, conv_tol=1e-6):
        y, mu = self.y, self.mu
        r = self.nb_r

        def d1(r):
            return float(np.sum(-(y + r) / (mu + r) - np.log(mu + r) + 1 + np.log(r) + digamma(r + y) - digamma(r)))

        def d2(r):
            return float(np.sum((y + r) / (mu + r) ** 2 - 2 / (mu + r) + 1 / r + polygamma(1, r + y) - polygamma(1, r)))

        def nb_logl(r):
            self.nb_r = r
            return self.loglikelihood()

        new_r, stepsize = 1.0, 1.0
        for _ in range(max_iter):
            dx, dx2 = d1(r), d2(r)
            increment = dx / dx2 if dx2 < 0 else dx
            new_r = r - stepsize * increment
            old_logl = nb_logl(r)
            for _ in range(20):
                if new_r <= 0:
                    stepsize /= 2
                    new_r = r - stepsize * increment
                else:
                    new_logl = nb_logl(new_r)
                    if old_logl >= new_logl:
                        stepsize /= 2
                        new_r = r - stepsize * increment
                    else:
                        break
            if abs(r - new_r) <= conv_tol:
                self.nb_r = new_r
                return new_r
            r = new_r
        self.nb_r = r
        return r


def iht_one_step(v: IHTVariable, old_logl: float, nstep: int):
    """`iht_one_step!` (src/fit.jl:213-263)."""
    eta = v.iht_stepsize()
    v.iht_gradstep(eta)
    v.update_xb(); v.update_mu()
    if v.est_r != "None":
        v.mle_for_r()
    new_logl = v.loglikelihood()
    eta_step = 0
    while (old_logl > new_logl) and (eta_step < nstep):      # `_iht_backtrack_` :484-486
        eta /= 2
        new_logl = v.backtrack(eta)
        eta_step += 1
    v.score()
    if np.isnan(new_logl):
        raise FloatingPointError("Loglikelihood function is NaN, aborting...")
    if np.isinf(new_logl):
        raise FloatingPointError("Loglikelihood function is Inf, aborting...")
    return eta, eta_step, new_logl


def fit_iht_loop(v: IHTVariable, tol=1e-4, max_iter=200, min_iter=5, max_step=3, trace: IHTTrace = None,
                 debias: bool = False):
    """`fit_iht!` (src/fit.jl:145-207).  Returns (best_logl, mm_iter)."""
    mm_iter = 0
    next_logl = -np.inf
    best_logl = -np.inf
    for it in range(1, max_iter + 1):
        if it >= max_iter:
            best_logl = v.save_prev(next_logl, best_logl)
            v.save_best_model()
            mm_iter = it
            break
        best_logl = v.save_prev(next_logl, best_logl)
        eta, eta_step, next_logl = iht_one_step(v, next_logl, max_step)
        if debias and it >= 5 and np.array_equal(v.idx, v.idx0):      # src/fit.jl:187-188
            v.debias()
        scaled_norm = v.check_convergence()
        if trace is not None:
            trace.logl.append(next_logl); trace.backtracks.append(eta_step)
            trace.tol.append(scaled_norm); trace.eta.append(eta)
            trace.support.append(np.flatnonzero(v.idx).copy())
        if it >= min_iter and scaled_norm < tol:
            best_logl = v.save_prev(next_logl, best_logl)
            v.save_best_model()
            mm_iter = it
            break
    return best_logl, mm_iter


def pve(y, mu):
    """`_pve` (src/pve.jl:22-24): var(mu)/var(y), sample variances."""
    return float(np.var(mu, ddof=1) / np.var(y, ddof=1))


def fit_iht(y, x, z=None, k=10, d=glm.NORMAL, l=None, zkeep=None, est_r="None", nb_r=1.0,
            tol=1e-4, max_iter=200, min_iter=5, max_step=3, cv_train_idx=None, init_beta=False,
            weight=None, debias=False, J=1, group=None) -> IHTResult:
    """`fit_iht` (src/fit.jl:60-118) on an oracle SnpLinAlg `x` (see oracle/snp.py)."""
    if z is None:
        z = np.ones(x.shape[0])
    if l is None:
        l = glm.IDENTITY
    if max_iter < 0 or max_step < 0 or J < 0:
        raise AssertionError("J, max_iter and max_step must be nonnegative")
    check_group(k, group)
    if not tol > np.finfo(np.float64).eps:
        raise AssertionError("Value of global tol must exceed machine precision!")
    v = IHTVariable(x, z, y, k, d, l, zkeep=zkeep, est_r=est_r, nb_r=nb_r, weight=weight, J=J, group=group)
    v.init_iht_indices(np.ones(v.n, bool) if cv_train_idx is None else cv_train_idx, init_beta)
    trace = IHTTrace()
    best_logl, mm_iter = fit_iht_loop(v, tol=tol, max_iter=max_iter, min_iter=min_iter,
                                      max_step=max_step, trace=trace, debias=debias)
    res = IHTResult(0.0, best_logl, mm_iter, v.best_b.copy(), v.best_c.copy(), J, k, [] if group is None else list(group), d,
                    pve(v.y, v.mu), trace, v.nb_r)
    res.v = v
    return res
