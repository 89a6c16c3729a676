"""Oracle parity at the SIZES BASELINE.json names (round-1 verdict: every oracle comparison was at n, p <= 3000).

Each case gives the CUDA path (through the C ABI) and the C+OpenMP restatement (oracle.cpu.PackedSnpLinAlgCPU under
oracle.iht / oracle.mviht / oracle.cv) the same host .bed bytes, y and z, and asserts the north_star bar: identical
support, iteration count and per-iteration backtracks; beta, c, loglikelihood (every iteration) and CV losses within
1e-6 relative.  The CPU side dominates the run time (about 40 s + 4 min + 1 min on 16 cores); IHTB_SCALE_TESTS=0 skips
the module, IHTB_SCALE_CV_FITS trims the configs[2] slice."""
import os

import numpy as np
import pytest

pytestmark = [pytest.mark.gpu, pytest.mark.scale]

import mendeliht_jl_b200 as m
from mendeliht_jl_b200 import synth
from oracle import cpu as ocpu
from oracle import cv as ocv
from oracle import glm, iht, mviht
from parity_helpers import RTOL, compare_fit, compare_mv_fit

if os.environ.get("IHTB_SCALE_TESTS", "1") == "0":
    pytest.skip("IHTB_SCALE_TESTS=0", allow_module_level=True)


def _operators(seed, n, p):
    bed = ocpu.synth_columns(seed, n, 0, p)          # the oracle's generator (twin of the device one, tested apart)
    return bed, m.B200SnpLinAlg.from_bed_columns(bed, n), ocpu.PackedSnpLinAlgCPU(bed, n)


def test_config1_bernoulli_50k_x_500k_matches_oracle():
    """BASELINE configs[1] (the bench workload, same seeds as bench.py): n=50k, p=500k, Bernoulli/Logit, k=20."""
    n, p, k = 50_000, 500_000, 20
    bed, g, o = _operators(2024, n, p)
    y, z, *_ = synth.simulate_response(2025, n, p, k, "Bernoulli", geno_seed=2024)
    res = m.fit_iht(y, g, z, k=k, d="Bernoulli", l="LogitLink")
    ref = iht.fit_iht(y, o, z, k=k, d=glm.BERNOULLI, l=glm.LOGIT)
    assert ref.iter < 200
    compare_fit(res, ref)
    # the decoded operator itself at this size: statistics bit for bit, one exact sweep against the CPU sweep
    mu, sinv, _ = g.stats()
    assert np.array_equal(mu, o.mu) and np.array_equal(sinv, o.sigma_inv)
    v = np.random.default_rng(1).normal(size=n)
    np.testing.assert_allclose(g.xt_v(v, m.SWEEP_EXACT), o.xt_v(v), rtol=1e-9, atol=1e-9)
    g.close()


def test_config2_poisson_cv_slice_100k_x_500k_matches_oracle():
    """BASELINE configs[2]: Poisson, q=5 folds x path 1:20 at n=100k, p=500k; an 8-fit slice of the 100-fit grid
    (folds 1 and 4, k in 3/6/9/12) against the oracle, fit by fit."""
    n, p, q = 100_000, 500_000, 5
    path = list(range(1, 21))
    bed, g, o = _operators(2025, n, p)
    y, z, *_ = synth.simulate_response(2025, n, p, 10, "Poisson", geno_seed=2025)
    folds = synth.folds_for(2025, n, q)
    grid = m.allocate_fold_and_k(q, path)
    combos = [i for i, (f, k) in enumerate(grid) if f in (1, 4) and k in (3, 6, 9, 12)]
    combos = combos[: int(os.environ.get("IHTB_SCALE_CV_FITS", "8"))]
    mses, iters = m.cv_iht(y, g, z, d="Poisson", l="LogLink", path=path, q=q, folds=folds, combos=combos,
                           return_grid=True)
    _, rm, ri = ocv.cv_iht(y, o, z, d=glm.POISSON, l=glm.LOG, path=path, q=q, folds=folds, combos_todo=set(combos),
                           return_grid=True)
    assert np.array_equal(iters[combos], ri[combos]), (iters[combos], ri[combos])
    np.testing.assert_allclose(mses[combos], rm[combos], rtol=RTOL)
    # the whole 100-fit grid in ONE library call (ihtb_cv_run: work queue, two fits at a time sharing PAIR sweeps whose
    # candidates are re-scored exactly) must reproduce the fit-by-fit loop bit for bit on the slice the oracle checked
    gm, gi = m.cv_run(y, g, z, folds, q, path, d="Poisson", l="LogLink")
    assert np.array_equal(gi[combos], iters[combos]) and np.array_equal(gm[combos], mses[combos])
    assert np.all(gi > 0) and np.all(gm > 0)
    g.close()


def test_config3_mvnormal_r5_100k_x_500k_matches_oracle():
    """BASELINE configs[3]: MvNormal, r=5 traits, n=100k, p=500k, k=50 (same generator as scripts/run_configs.py mv)."""
    n, p, r, k = 100_000, 500_000, 5, 50
    bed, g, o = _operators(2026, n, p)
    rng = np.random.default_rng(2026)
    idx = np.sort(rng.permutation(p)[:k])
    B = np.zeros((r, k))
    for c in range(k):
        B[rng.integers(0, r), c] = rng.normal()
    A = rng.normal(size=(r, r)); cov = A @ A.T / r + 0.5 * np.eye(r)
    Y = np.linalg.cholesky(cov) @ rng.normal(size=(r, n)) + 1.0
    Y += B @ o.columns(idx).T
    ref = mviht.fit_mv_iht(Y, o, None, k=k)
    # FAST: one single-vector pass per trait; PAIR: the skinny X'R, one pass over the matrix per two traits
    for mode in (m.SWEEP_FAST, m.SWEEP_PAIR):
        res = m.fit_iht(Y, g, None, k=k, sweep_mode=mode)
        compare_mv_fit(res, ref)
    g.close()
