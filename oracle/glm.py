"""Oracle (test infrastructure): GLM link / variance / deviance / logpdf formulas.

MendelIHT imports these from GLM.jl 1.x and Distributions.jl 0.25 (`src/MendelIHT.jl:7`;
call sites `src/utilities.jl:32-43,56,80,130,402,749`).  Neither package is vendored in
/root/reference, so their published formulas are restated here.
"""
from __future__ import annotations

import numpy as np
from scipy.special import gammaln, erf, xlogy, betaln, ndtri

NORMAL, BERNOULLI, POISSON, NEGBIN = "Normal", "Bernoulli", "Poisson", "NegativeBinomial"
IDENTITY, LOGIT, LOG, PROBIT, CLOGLOG, CAUCHIT, SQRT, INVERSE, INVSQ = (
    "IdentityLink", "LogitLink", "LogLink", "ProbitLink", "CloglogLink",
    "CauchitLink", "SqrtLink", "InverseLink", "InverseSquareLink")

DIST_ID = {NORMAL: 0, BERNOULLI: 1, POISSON: 2, NEGBIN: 3}
LINK_ID = {IDENTITY: 0, LOGIT: 1, LOG: 2, PROBIT: 3, CLOGLOG: 4, CAUCHIT: 5, SQRT: 6,
           INVERSE: 7, INVSQ: 8}


def canonicallink(d: str) -> str:
    """`canonicallink(d())`, except LogLink for NegativeBinomial (reference `src/wrapper.jl:87`)."""
    return {NORMAL: IDENTITY, BERNOULLI: LOGIT, POISSON: LOG, NEGBIN: LOG}[d]


def linkinv(l: str, eta):
    eta = np.asarray(eta, dtype=np.float64)
    if l == IDENTITY:
        return eta.copy()
    if l == LOGIT:
        return 1.0 / (1.0 + np.exp(-eta))
    if l == LOG:
        return np.exp(eta)
    if l == PROBIT:
        return 0.5 * (1.0 + erf(eta / np.sqrt(2.0)))
    if l == CLOGLOG:
        return -np.expm1(-np.exp(eta))
    if l == CAUCHIT:
        return 0.5 + np.arctan(eta) / np.pi
    if l == SQRT:
        return eta * eta
    if l == INVERSE:
        return 1.0 / eta
    if l == INVSQ:
        return 1.0 / np.sqrt(eta)
    raise ValueError(l)


def mueta(l: str, eta):
    eta = np.asarray(eta, dtype=np.float64)
    if l == IDENTITY:
        return np.ones_like(eta)
    if l == LOGIT:
        e = np.exp(-np.abs(eta))
        f = 1.0 + e
        return e / (f * f)
    if l == LOG:
        return np.exp(eta)
    if l == PROBIT:
        return np.exp(-0.5 * eta * eta) / np.sqrt(2.0 * np.pi)
    if l == CLOGLOG:
        return np.exp(eta) * np.exp(-np.exp(eta))
    if l == CAUCHIT:
        return 1.0 / (np.pi * (1.0 + eta * eta))
    if l == SQRT:
        return 2.0 * eta
    if l == INVERSE:
        return -1.0 / (eta * eta)
    if l == INVSQ:
        mu = 1.0 / np.sqrt(eta)
        return -(mu ** 3) / 2.0
    raise ValueError(l)


def glmvar(d: str, mu, r: float = 1.0):
    if d == NORMAL:
        return np.ones_like(mu)
    if d == BERNOULLI:
        return mu * (1.0 - mu)
    if d == POISSON:
        return mu
    if d == NEGBIN:
        return mu * (1.0 + mu / r)
    raise ValueError(d)


def devresid(d: str, y, mu, r: float = 1.0):
    with np.errstate(divide="ignore", invalid="ignore"):
        if d == NORMAL:
            return (y - mu) ** 2
        if d == BERNOULLI:
            one = -2.0 * np.log(mu)
            zero = -2.0 * np.log1p(-mu)
            gen = 2.0 * (xlogy(y, y / mu) + xlogy(1.0 - y, (1.0 - y) / (1.0 - mu)))
            return np.where(y == 1, one, np.where(y == 0, zero, gen))
        if d == POISSON:
            return 2.0 * (xlogy(y, y / mu) - (y - mu))
        if d == NEGBIN:
            v = 2.0 * (xlogy(y, y / mu) + xlogy(y + r, (mu + r) / (y + r)))
            return np.where(mu == 0, np.nan, v)
    raise ValueError(d)


def logpdf(d: str, y, mu, phi: float, r: float = 1.0):
    """Per-observation log density as `loglik_obs` builds it (reference `src/utilities.jl:32-43`)."""
    with np.errstate(divide="ignore", invalid="ignore"):
        if d == NORMAL:
            sigma = np.sqrt(phi)
            z = (y - mu) / sigma
            return -(z * z + np.log(2.0 * np.pi)) / 2.0 - np.log(sigma)
        if d == BERNOULLI:
            return np.where(y == 1, np.log(mu), np.log(1.0 - mu))
        if d == POISSON:
            return xlogy(y, mu) - mu - gammaln(y + 1.0)
        if d == NEGBIN:
            p = r / (mu + r)
            return r * np.log(p) + y * np.log1p(-p) - np.log(y + r) - betaln(r, y + 1.0)
    raise ValueError(d)


def deviance(d: str, y, mu, wts, r: float = 1.0) -> float:
    """`deviance(d, y, mu, wts)` (reference `src/utilities.jl:52-59`): sequential weighted sum."""
    return float(np.sum(wts * devresid(d, y, mu, r)))


def loglikelihood(d: str, y, mu, wts, r: float = 1.0) -> float:
    """`loglikelihood(v)` (reference `src/utilities.jl:9-20`): phi = deviance / length(y) even under
    CV masks."""
    phi = deviance(d, y, mu, wts, r) / y.shape[0]
    return float(np.sum(wts * logpdf(d, y, mu, phi, r)))


# ---- GLM.jl `fit(GeneralizedLinearModel, X, y, d, l)` (used by `debias!`, reference src/utilities.jl:1014-1020) --------
def linkfun(l: str, mu):
    """GLM.jl `linkfun` (the inverse of `linkinv`)."""
    mu = np.asarray(mu, dtype=np.float64)
    if l == IDENTITY:
        return mu.copy()
    if l == LOGIT:
        return np.log(mu / (1.0 - mu))
    if l == LOG:
        return np.log(mu)
    if l == PROBIT:
        return ndtri(mu)
    if l == CLOGLOG:
        return np.log(-np.log1p(-mu))
    if l == CAUCHIT:
        return np.tan(np.pi * (mu - 0.5))
    if l == SQRT:
        return np.sqrt(mu)
    if l == INVERSE:
        return 1.0 / mu
    if l == INVSQ:
        return 1.0 / (mu * mu)
    raise ValueError(l)


def mustart(d: str, y):
    """GLM.jl `mustart(d, y, wt=1)`: the starting mean of IRLS."""
    y = np.asarray(y, dtype=np.float64)
    if d == NORMAL:
        return y.copy()
    if d == BERNOULLI:
        return (y + 0.5) / 2.0
    if d == POISSON:
        return y + 0.1
    if d == NEGBIN:
        return np.where(y == 0, y + 1.0 / 6.0, y)
    raise ValueError(d)


class ConvergenceException(RuntimeError):
    pass


def glm_fit(X, y, d: str, l: str, r: float = 1.0, maxiter=30, minstepfac=0.001, atol=1e-6, rtol=1e-6):
    """IRLS exactly as GLM.jl 1.x `_fit!` runs it with default arguments (no weights, no offset, no intercept added):
    start from eta = linkfun(mustart(y)), one weighted least-squares solve on the working response, then Newton steps
    with step-halving while the deviance increases, stop when devold - dev < max(rtol*devold, atol).
    Working weights are mueta^2/var (GLM.jl uses the algebraically equal mueta for canonical links)."""
    X = np.asarray(X, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)

    def update(eta):
        mu = linkinv(l, eta)
        dmu = mueta(l, eta)
        wrkres = (y - mu) / dmu
        wrkwt = dmu * dmu / glmvar(d, mu, r)
        return wrkres, wrkwt, float(np.sum(devresid(d, y, mu, r)))

    def delbeta(resp, wt):
        xw = X * wt[:, None]
        return np.linalg.solve(xw.T @ X, xw.T @ resp)

    eta = linkfun(l, mustart(d, y))
    wrkres, wrkwt, _ = update(eta)
    beta0 = delbeta(wrkres + eta, wrkwt)
    wrkres, wrkwt, devold = update(X @ beta0)
    for _ in range(maxiter):
        f = 1.0
        delta = delbeta(wrkres, wrkwt)
        wrkres, wrkwt, dev = update(X @ (beta0 + delta))
        if np.isnan(dev):
            dev = np.inf
        while dev > devold + rtol * dev:
            f /= 2.0
            if not f > minstepfac:
                raise RuntimeError(f"step-halving failed at beta0 = {beta0}")
            wrkres, wrkwt, dev = update(X @ (beta0 + f * delta))
            if np.isnan(dev):
                dev = np.inf
        beta0 = beta0 + f * delta
        if devold - dev < max(rtol * devold, atol):
            return beta0
        devold = dev
    raise ConvergenceException(f"failure to converge after {maxiter} iterations.")
