"""Oracle (test infrastructure): multivariate-Normal IHT, numpy restatement of the reference.

Follows /root/reference `src/multivariate.jl` function by function:
  `loglikelihood` :9-13, `update_xb!` :21-31, `update_mu!` :39-43, `update_resid!` :50-58,
  `score!`/`update_df!` :66-92, `_iht_gradstep!`/`project_k!` :99-127, `vectorize!` :138-161,
  `update_support!` :197-206, `iht_stepsize!` :220-254, `solve_Sigma!` :276-282, `_choose!` :310-351,
  `save_prev!` :356-367, `init_iht_indices!` :376-452, `check_convergence` :454-458,
  `backtrack!` :460-473, `save_best_model!` :485-496; driver `src/fit.jl:145-263`; `pve` src/pve.jl:35-37.

Layout as in the reference: Y is r x n, X is p x n (Transpose(SnpLinAlg)), Z is q x n, B is r x p.

Restrictions (documented): `zkeep` must be all-true.  With a `false` entry the reference's matrix
`unvectorize!` (:172-189) reads the wrong slice of `full_b` (its cursor skips kept columns that
`vectorize!` did emit), so there is no well-defined behaviour to reproduce.
Quirk reproduced: `iht_stepsize!` overwrites Gamma with its *pivoted* Cholesky factor and ignores the
permutation (:241-246).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from .iht import project_k


def pivoted_cholesky_upper(a: np.ndarray) -> np.ndarray:
    """U with P'AP = U'U as LAPACK dpstrf(uplo='U') orders it (max remaining diagonal first,
    first maximum wins); the permutation is dropped, as `iht_stepsize!` does."""
    s = np.array(a, dtype=np.float64, copy=True)
    s = np.triu(s) + np.triu(s, 1).T          # Symmetric(A, :U)
    n = s.shape[0]
    u = np.zeros_like(s)
    for j in range(n):
        d = np.diag(s)[j:]
        pvt = j + int(np.argmax(d))
        if pvt != j:
            s[[j, pvt], :] = s[[pvt, j], :]
            s[:, [j, pvt]] = s[:, [pvt, j]]
            u[:, [j, pvt]] = u[:, [pvt, j]]
        ajj = s[j, j]
        if not ajj > 0:
            raise np.linalg.LinAlgError("RankDeficientException")
        ajj = np.sqrt(ajj)
        u[j, j] = ajj
        u[j, j + 1:] = s[j, j + 1:] / ajj
        s[j + 1:, j + 1:] -= np.outer(u[j, j + 1:], u[j, j + 1:])
    return u


@dataclass
class MvTrace:
    logl: list = field(default_factory=list)
    backtracks: list = field(default_factory=list)
    tol: list = field(default_factory=list)
    eta: list = field(default_factory=list)


@dataclass
class MvIHTResult:
    """`mIHTResult` (src/data_structures.jl:263-275)."""
    time: float
    logl: float
    iter: int
    beta: np.ndarray      # r x p
    c: np.ndarray         # r x q
    k: int
    traits: int
    Sigma: np.ndarray
    sigma_g: np.ndarray
    trace: MvTrace = None


class MvIHTVariable:
    def __init__(self, x, z, y, k, zkeep=None):
        """x: oracle SnpLinAlg (n x p; used as its transpose), y: r x n, z: q x n."""
        self.x = x
        self.Y = np.asarray(y, dtype=np.float64)
        self.Z = np.asarray(z, dtype=np.float64)
        n, p = x.shape
        self.r, self.q = self.Y.shape[0], self.Z.shape[0]
        if not (n == self.Y.shape[1] == self.Z.shape[1]):
            raise ValueError(f"number of samples in y, x, and z = {self.Y.shape[1]}, {n}, {self.Z.shape[1]} are not equal")
        self.n, self.p, self.k = n, p, int(k)
        self.zkeep = np.ones(self.q, bool) if zkeep is None else np.asarray(zkeep, bool)
        if self.zkeep.shape[0] != self.q:
            raise ValueError(f"zkeep must have length {self.q} but was {self.zkeep.shape[0]}")
        if not self.zkeep.all():
            raise NotImplementedError("multivariate zkeep with false entries is ill-defined in the reference")
        self.zkeepn = self.r * int(self.zkeep.sum())

    def nsamples(self):
        return int(np.count_nonzero(self.cv_wts))

    # :376-452
    # :519-558 (`initialize_beta!`, executed in trait order; the reference adds the intercepts to a shared `C0` from
    # several threads without synchronisation, the restatement is the single-thread result)
    def initialize_beta(self, cv_idx):
        from .iht import linreg_columns
        cv = np.asarray(cv_idx, dtype=bool)
        for j in range(self.r):
            ys = self.Y[j, cv]
            c0 = 0.0
            for j0 in range(0, self.p, 2048):
                xs = self.x.columns(np.arange(j0, min(j0 + 2048, self.p)))[cv]
                icpt, slope = linreg_columns(xs, ys)
                c0 += float(icpt.sum())
                self.B[j, j0:j0 + 2048] = slope
            if self.q > 1:
                icpt, slope = linreg_columns(self.Z[1:, cv].T, ys)
                c0 += float(icpt.sum())
                self.C[j, 1:] = slope
            self.C[j, 0] = c0 / (self.p + self.q - 1)
        np.clip(self.C, -2, 2, out=self.C)
        np.clip(self.B, -2, 2, out=self.B)
        self.B0 = self.B.copy(); self.C0 = self.C.copy()

    def init_iht_indices(self, cv_idx, init_beta=False):
        if self.k < 1:
            raise ValueError("Multivariate IHT requires k >= 1!")
        r, p, q, n = self.r, self.p, self.q, self.n
        self.B = np.zeros((r, p)); self.B0 = np.zeros((r, p)); self.best_B = np.zeros((r, p))
        self.C = np.zeros((r, q)); self.C0 = np.zeros((r, q)); self.best_C = np.zeros((r, q))
        self.BX = np.zeros((r, n))
        self.idx = np.zeros(p, bool); self.idx0 = np.zeros(p, bool)
        self.idc = self.zkeep.copy(); self.idc0 = self.zkeep.copy()
        self.Gamma = np.eye(r); self.Gamma0 = np.eye(r)
        self.cv_wts = np.zeros(n); self.cv_wts[np.asarray(cv_idx, bool)] = 1.0
        nz = self.nsamples()
        self.C[:, 0] = (self.Y * self.cv_wts[None, :]).sum(axis=1) / nz
        self.CZ = self.C @ self.Z
        if init_beta:                              # :425-429
            self.initialize_beta(cv_idx)
            self._project()
            self.update_xb()
        self.update_mu(); self.update_resid(); self.score()
        if init_beta:
            return
        full = self._vectorize(self.df, self.df2)
        project_k(full, self.k + self.zkeepn)
        self._unvectorize(full, self.df, self.df2)
        self.idx = (self.df != 0).any(axis=0)
        self.idc = (self.df2 != 0).any(axis=0)

    # :138-189 (all-kept covariates only)
    def _vectorize(self, B, C):
        return np.concatenate([B.reshape(-1, order="F"), np.full(C.size, np.inf)])

    def _unvectorize(self, a, B, C):
        B[...] = a[: B.size].reshape(B.shape, order="F")

    # :21-31 (memory-efficient branch: getindex path on the support rows)
    def update_xb(self):
        idx = np.flatnonzero(self.idx)
        xs = self.x.columns(idx)               # n x |idx|
        self.BX = self.B[:, idx] @ xs.T
        self.CZ = self.C @ self.Z

    def update_mu(self):
        self.mu = self.BX + self.CZ

    def update_resid(self):
        self.resid = (self.Y - self.mu) * self.cv_wts[None, :]

    # :66-92
    def score(self):
        r_by_n1 = self.Gamma @ self.resid
        self.df = self.x.xt_v(np.ascontiguousarray(r_by_n1.T)).T     # p x r -> r x p
        self.df2 = r_by_n1 @ self.Z.T

    # :220-254
    def iht_stepsize(self):
        idx = np.flatnonzero(self.idx)
        dfidx = self.df[:, idx]
        numer = float(np.sum(dfidx ** 2))
        v = (dfidx @ self.x.columns(idx).T) * self.cv_wts[None, :]
        self.Gamma = pivoted_cholesky_upper(self.Gamma)
        uv = self.Gamma @ v
        denom = float(np.sum(uv ** 2))
        with np.errstate(divide="ignore", invalid="ignore"):
            eta = numer / denom if denom != 0 else (np.inf if numer > 0 else np.nan)
        if np.isinf(eta) or np.isnan(eta):
            eta = 1e-8
        return eta

    # :99-127, :310-351
    def iht_gradstep(self, eta):
        self.B += eta * self.df
        self.C += eta * self.df2
        self._project()

    # project_k!(v::mIHTVariable) :106-127
    def _project(self):
        full = self._vectorize(self.B, self.C)
        project_k(full, self.k + self.zkeepn)
        self._unvectorize(full, self.B, self.C)
        # `_choose!`: excess counted against k + zkeepn with only B (and non-kept C) entries
        b_nz = int(np.count_nonzero(self.B))
        excess = b_nz - (self.k + self.zkeepn)
        if excess > 0:
            flat = self.B.reshape(-1, order="F")
            pos = np.flatnonzero(flat)
            order = sorted(pos, key=lambda j: (abs(flat[j]), -j))
            flat[order[:excess]] = 0.0
            self.B = flat.reshape(self.B.shape, order="F")
        self.idx = (self.B != 0).any(axis=0)
        self.idc = (self.C != 0).any(axis=0)

    # :276-282
    def solve_sigma(self):
        self.update_resid()
        s = self.resid @ self.resid.T / self.nsamples()
        l = np.linalg.cholesky(s)
        linv = np.linalg.inv(l)
        self.Gamma = linv.T @ linv

    # :9-13
    def loglikelihood(self):
        rr = self.resid @ self.resid.T
        sign, logdet = np.linalg.slogdet(self.Gamma)
        if sign <= 0:
            return np.nan
        return self.nsamples() / 2 * logdet - 0.5 * float(np.trace(self.Gamma @ rr))

    def save_prev(self, cur_logl, best_logl):
        self.B0 = self.B.copy(); self.C0 = self.C.copy()
        self.idx0 = self.idx.copy(); self.idc0 = self.idc.copy(); self.Gamma0 = self.Gamma.copy()
        if cur_logl > best_logl:
            self.best_B = self.B.copy(); self.best_C = self.C.copy()
        return max(cur_logl, best_logl)

    def check_convergence(self):
        the_norm = max(np.max(np.abs(self.B - self.B0)), np.max(np.abs(self.C - self.C0)))
        return the_norm / (max(np.max(np.abs(self.B0)), np.max(np.abs(self.C0))) + 1.0)

    def backtrack(self, eta):
        self.B = self.B0.copy(); self.C = self.C0.copy(); self.Gamma = self.Gamma0.copy()
        self.iht_gradstep(eta)
        self.update_xb(); self.update_mu(); self.solve_sigma()
        return self.loglikelihood()

    def save_best_model(self):
        self.B = self.best_B.copy(); self.C = self.best_C.copy()
        self.idx = (self.B != 0).any(axis=0); self.idc = (self.C != 0).any(axis=0)
        self.update_xb(); self.update_mu()


def mv_one_step(v: MvIHTVariable, old_logl, nstep):
    """`iht_one_step!` (src/fit.jl:213-263), mIHTVariable branch."""
    eta = v.iht_stepsize()
    v.iht_gradstep(eta)
    v.update_xb(); v.update_mu(); v.solve_sigma()
    new_logl = v.loglikelihood()
    eta_step = 0
    while (old_logl > new_logl) and (eta_step < nstep):
        eta /= 2
        new_logl = v.backtrack(eta)
        eta_step += 1
    v.score()
    if np.isnan(new_logl):
        raise FloatingPointError("Loglikelihood function is NaN, aborting...")
    if np.isinf(new_logl):
        raise FloatingPointError("Loglikelihood function is Inf, aborting...")
    return eta, eta_step, new_logl


def mv_fit_loop(v, tol=1e-4, max_iter=200, min_iter=5, max_step=3, trace=None):
    mm_iter, next_logl, best_logl = 0, -np.inf, -np.inf
    for it in range(1, max_iter + 1):
        if it >= max_iter:
            best_logl = v.save_prev(next_logl, best_logl); v.save_best_model(); mm_iter = it
            break
        best_logl = v.save_prev(next_logl, best_logl)
        eta, eta_step, next_logl = mv_one_step(v, next_logl, max_step)
        scaled = v.check_convergence()
        if trace is not None:
            trace.logl.append(next_logl); trace.backtracks.append(eta_step)
            trace.tol.append(scaled); trace.eta.append(eta)
        if it >= min_iter and scaled < tol:
            best_logl = v.save_prev(next_logl, best_logl); v.save_best_model(); mm_iter = it
            break
    return best_logl, mm_iter


def fit_mv_iht(Y, x, Z=None, k=10, zkeep=None, tol=1e-4, max_iter=200, min_iter=5, max_step=3,
               cv_train_idx=None, init_beta=False) -> MvIHTResult:
    """`fit_iht(Y, Transpose(xla), Z; k)` for MvNormal (src/fit.jl:60-118)."""
    Y = np.asarray(Y, dtype=np.float64)
    if Z is None:
        Z = np.ones((1, Y.shape[1]))
    v = MvIHTVariable(x, Z, Y, k, zkeep)
    v.init_iht_indices(np.ones(v.n, bool) if cv_train_idx is None else cv_train_idx, init_beta)
    trace = MvTrace()
    best_logl, mm_iter = mv_fit_loop(v, tol, max_iter, min_iter, max_step, trace)
    sg = np.array([np.var(v.mu[i], ddof=1) / np.var(v.Y[i], ddof=1) for i in range(v.r)])
    res = MvIHTResult(0.0, best_logl, mm_iter, v.best_B.copy(), v.best_C.copy(), k, v.r,
                      np.linalg.inv(v.Gamma), sg, trace)
    res.v = v
    return res
