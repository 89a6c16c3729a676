"""Oracle (test infrastructure): GLM link / variance / deviance / logpdf formulas.

MendelIHT imports these from GLM.jl 1.x and Distributions.jl 0.25 (`src/MendelIHT.jl:7`;
call sites `src/utilities.jl:32-43,56,80,130,402,749`).  Neither package is vendored in
/root/reference, so their published formulas are restated here.
"""
from __future__ import annotations

import numpy as np
from scipy.special import gammaln, erf, xlogy, betaln

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
