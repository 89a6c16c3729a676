"""Oracle (test infrastructure): q-fold x k-path cross-validation driver, numpy restatement.

Follows /root/reference `src/cross_validation.jl`: `cv_iht` :60-131, `allocate_fold_and_k` :217-223,
`predict!` :279-299, `meanloss` :304-320.  Folds must be passed explicitly (the reference draws them
from the global RNG, :72).  Every fit sweeps all n rows with masks; mu_j/sigma_j stay full-sample.
"""
from __future__ import annotations

import numpy as np

from . import glm
from .iht import IHTVariable, fit_iht_loop
from .mviht import MvIHTVariable, mv_fit_loop


def allocate_fold_and_k(q: int, path):
    """Fold-major list of (fold, k), folds 1..q (:217-223)."""
    return [(fold, int(k)) for fold in range(1, q + 1) for k in path]


def meanloss(fitloss, q: int, folds):
    """Fold-size weighted sum over folds (:304-320)."""
    folds = np.asarray(folds)
    ninfold = np.array([(folds == f).sum() for f in range(1, q + 1)])
    pathsize = len(fitloss) // q
    loss = np.zeros(pathsize)
    for j in range(q):
        wfold = ninfold[j] / folds.shape[0]
        for i in range(pathsize):
            loss[i] += fitloss[i + j * pathsize] * wfold
    return loss


def cv_iht(y, x, z=None, d=glm.NORMAL, l=glm.IDENTITY, path=range(1, 21), q=5, folds=None,
           zkeep=None, max_iter=100, min_iter=5, nb_r=1.0, return_grid=False, init_beta=False, weight=None,
           debias=False, J=1, group=None, combos_todo=None):
    """`cv_iht` (:60-131), univariate or multivariate by the shape of y.  `combos_todo`: optional subset of grid
    positions to run (slices of grids too large to run whole on the CPU); the others stay 0."""
    y = np.asarray(y, dtype=np.float64)
    multivariate = y.ndim == 2 and y.shape[0] > 1 and y.shape[1] > 1
    n = x.shape[0]
    if folds is None:
        raise ValueError("oracle cv_iht needs explicit folds")
    folds = np.asarray(folds)
    path = [int(k) for k in path]
    if max(path) > x.shape[1]:
        raise ValueError("Sparsity level in `path` cannot be larger than total number of variables")
    if z is None:
        z = np.ones((1, n)) if multivariate else np.ones(n)
    combos = allocate_fold_and_k(q, path)
    mses = np.zeros(len(combos))
    iters = np.zeros(len(combos), dtype=np.int64)
    for i, (fold, k) in enumerate(combos):
        if combos_todo is not None and i not in combos_todo:
            continue
        test = folds == fold
        train = ~test
        if multivariate:
            v = MvIHTVariable(x, z, y, k, zkeep)
            v.init_iht_indices(train, init_beta)
            _, iters[i] = mv_fit_loop(v, max_iter=max_iter, min_iter=min_iter)
            v.cv_wts[train] = 0.0; v.cv_wts[test] = 1.0
            v.update_xb(); v.update_mu()
            mses[i] = float(np.sum((v.Y - v.mu) ** 2 * v.cv_wts[None, :]))
        else:
            v = IHTVariable(x, z, y, k, d, l, zkeep=zkeep, nb_r=nb_r, weight=weight, J=J, group=group)
            v.init_iht_indices(train, init_beta)
            _, iters[i] = fit_iht_loop(v, max_iter=max_iter, min_iter=min_iter, debias=debias)
            v.cv_wts[train] = 0.0; v.cv_wts[test] = 1.0
            v.update_xb(); v.update_mu()
            mses[i] = v.deviance()
    mse = meanloss(mses, q, folds)
    if return_grid:
        return mse, mses, iters
    return mse
