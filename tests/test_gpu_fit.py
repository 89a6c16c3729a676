"""GPU parity of the IHT loop against the CPU oracle and the reference's published trace, through the C ABI.

Bar (BASELINE.json north_star): support indices and iteration counts identical; beta, loglikelihood and CV MSEs
within 1e-6 relative (the fit re-scores every top-k candidate in FP64, so this holds for both sweep modes)."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import mendeliht_jl_b200 as m
from mendeliht_jl_b200 import synth
from oracle import cv as ocv
from oracle import glm, iht, snp
from conftest import GOLDEN
from parity_helpers import RTOL, compare_fit as _compare, compare_mv_fit as _compare_mv

MODES = [m.SWEEP_FAST, m.SWEEP_EXACT]


@pytest.mark.parametrize("mode", MODES)
def test_docs_trace_normal_k7(normal_data, mode):
    """docs/src/man/examples.md:230-268 straight against the CUDA path."""
    gold = json.load(open(os.path.join(GOLDEN, "docs_trace_normal_k7.json")))
    g = m.B200SnpLinAlg.from_bed_columns(normal_data["bed"], normal_data["n"])
    res = m.fit_iht(normal_data["y"], g, normal_data["z"], k=7, d="Normal", l="IdentityLink", sweep_mode=mode)
    assert res.iter == gold["iter"]
    np.testing.assert_allclose([t[0] for t in res.trace], gold["logl"], rtol=1e-9)
    np.testing.assert_allclose([t[2] for t in res.trace], gold["tol"], rtol=1e-6)
    assert [t[1] for t in res.trace] == gold["backtracks"]
    assert list(np.flatnonzero(res.beta) + 1) == gold["support_1based"]
    np.testing.assert_allclose(res.beta[np.flatnonzero(res.beta)], gold["beta_6sig"], rtol=2e-6)
    np.testing.assert_allclose(res.c, gold["c_6sig"], rtol=2e-6)
    assert abs(res.sigma_g - gold["pve"]) < 1e-9
    assert abs(res.logl - gold["final_logl"]) < 1e-7


@pytest.mark.parametrize("mode", MODES)
def test_config1_readme_call(normal_data, normal_oracle, mode):
    """BASELINE config 1: iht(datadir/normal, 9, Normal) (README.md:104) vs the oracle."""
    g = m.B200SnpLinAlg.from_bed_columns(normal_data["bed"], normal_data["n"])
    res = m.fit_iht(normal_data["y"], g, None, k=9, sweep_mode=mode)
    ref = iht.fit_iht(normal_data["y"], normal_oracle, None, k=9)
    _compare(res, ref)
    assert res.iter == 10 and list(np.flatnonzero(res.beta) + 1) == [1266, 3137, 4246, 4717, 6290, 7629, 7755, 8375, 9415]


CASES = [
    # d, link, n, p, true k, fitted k, covariates, missing rate
    ("Normal", "IdentityLink", 1200, 3000, 8, 10, 2, 0.0),
    ("Bernoulli", "LogitLink", 2000, 3000, 5, 7, 1, 0.0),
    ("Bernoulli", "LogitLink", 3000, 2000, 4, 4, 0, 0.0),
    ("Poisson", "LogLink", 1500, 2500, 6, 8, 1, 0.0),
    ("NegativeBinomial", "LogLink", 1500, 2500, 6, 8, 0, 0.0),
    ("Normal", "IdentityLink", 1003, 2001, 5, 7, 1, 0.01),      # ragged n, missing genotypes
    ("Bernoulli", "LogitLink", 1111, 1500, 4, 4, 2, 0.005),
]


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("d,l,n,p,k,kfit,ncov,miss", CASES)
def test_fit_matches_oracle(d, l, n, p, k, kfit, ncov, miss, mode):
    seed = 100 + n + p
    y, z, _, _, _ = synth.simulate_response(seed, n, p, k, d, n_cov=ncov, missing_rate=miss)
    bed = synth.packed_columns(seed, n, np.arange(p), miss)
    g = m.B200SnpLinAlg.from_bed_columns(bed, n)
    o = snp.SnpLinAlgOracle(bed, n)
    res = m.fit_iht(y, g, z, k=kfit, d=d, l=l, nb_r=10.0, sweep_mode=mode)
    ref = iht.fit_iht(y, o, z, k=kfit, d=d, l=l, nb_r=10.0)
    assert ref.iter < 200, "parity cases must converge (see test_oscillating_fit for the other kind)"
    _compare(res, ref)


@pytest.mark.parametrize("mode", MODES)
def test_oscillating_fit(mode):
    """A logistic fit with k larger than the signal never converges: it keeps taking max_step backtracks, accepts
    likelihood drops and pushes eta to the +-20 clamp.  There w = mueta/glmvar = e/(1+e)^2 / (mu (1-mu)) loses ~9
    digits to cancellation in 1-mu (the reference's own formula, src/utilities.jl:130), so ulp-level differences
    between libm implementations are amplified to ~1e-6 for a few iterations before the map contracts again.
    Support, iteration count and backtracks still agree exactly; values are compared at 1e-4."""
    d, l, n, p, k = "Bernoulli", "LogitLink", 1500, 3000, 6
    seed = 100 + n + p
    y, z, _, _, _ = synth.simulate_response(seed, n, p, k, d)
    bed = synth.packed_columns(seed, n, np.arange(p))
    g = m.B200SnpLinAlg.from_bed_columns(bed, n)
    res = m.fit_iht(y, g, z, k=k + 2, d=d, l=l, sweep_mode=mode)
    ref = iht.fit_iht(y, snp.SnpLinAlgOracle(bed, n), z, k=k + 2, d=d, l=l)
    assert ref.iter == 200
    _compare(res, ref, rtol=1e-4)


def test_zkeep_lets_covariates_compete():
    n, p, k = 1200, 2000, 5
    y, z, _, _, _ = synth.simulate_response(7, n, p, k, "Normal", n_cov=3)
    bed = synth.packed_columns(7, n, np.arange(p))
    g = m.B200SnpLinAlg.from_bed_columns(bed, n)
    o = snp.SnpLinAlgOracle(bed, n)
    zkeep = np.array([True, False, True, False])
    res = m.fit_iht(y, g, z, k=6, zkeep=zkeep)
    ref = iht.fit_iht(y, o, z, k=6, zkeep=zkeep)
    _compare(res, ref)
    assert np.count_nonzero(res.beta) + np.count_nonzero(res.c[~zkeep]) <= 6


def test_max_iter_exit_and_min_iter():
    n, p = 1000, 1500
    y, z, _, _, _ = synth.simulate_response(3, n, p, 5, "Normal")
    bed = synth.packed_columns(3, n, np.arange(p))
    g = m.B200SnpLinAlg.from_bed_columns(bed, n)
    o = snp.SnpLinAlgOracle(bed, n)
    for kw in ({"max_iter": 3}, {"max_iter": 1}, {"min_iter": 8}, {"tol": 1e-2}, {"max_step": 0}):
        res = m.fit_iht(y, g, z, k=5, **kw)
        ref = iht.fit_iht(y, o, z, k=5, **kw)
        _compare(res, ref)


def test_cv_matches_oracle():
    n, p = 1000, 2000
    y, z, _, _, _ = synth.simulate_response(21, n, p, 4, "Poisson", n_cov=1)
    bed = synth.packed_columns(21, n, np.arange(p))
    g = m.B200SnpLinAlg.from_bed_columns(bed, n)
    o = snp.SnpLinAlgOracle(bed, n)
    folds = synth.folds_for(21, n, 3)
    path = [1, 3, 5, 8]
    mses, iters = m.cv_iht(y, g, z, d="Poisson", l="LogLink", path=path, q=3, folds=folds, return_grid=True)
    rmse, rgrid, riters = ocv.cv_iht(y, o, z, d=glm.POISSON, l=glm.LOG, path=path, q=3, folds=folds, return_grid=True)
    assert np.array_equal(iters, riters)
    np.testing.assert_allclose(mses, rgrid, rtol=RTOL)
    np.testing.assert_allclose(m.meanloss(mses, 3, folds), rmse, rtol=RTOL)
    assert np.all(mses > 0)


def test_cv_run_single_call_matches_loop():
    """`ihtb_cv_run` (the whole grid in one library call) == the host loop over fit handles == the oracle."""
    n, p, k = 1500, 2500, 5
    y, z, _, _, _ = synth.simulate_response(23, n, p, k, "Poisson", n_cov=1, missing_rate=0.002)
    bed = synth.packed_columns(23, n, np.arange(p), 0.002)
    g = m.B200SnpLinAlg.from_bed_columns(bed, n)
    folds = synth.folds_for(23, n, 3)
    path = [2, 4, 7]
    mses, iters = m.cv_run(y, g, z, folds, 3, path, d="Poisson", l="LogLink")
    lm, li = m.cv_iht(y, g, z, d="Poisson", l="LogLink", path=path, q=3, folds=folds, return_grid=True)
    np.testing.assert_array_equal(mses, lm)
    np.testing.assert_array_equal(iters, li)
    _, rgrid, riters = ocv.cv_iht(y, snp.SnpLinAlgOracle(bed, n), z, d="Poisson", l="LogLink", path=path, q=3,
                                  folds=folds, return_grid=True)
    assert np.array_equal(iters, riters) and riters.max() < 100      # converging fits (cf. test_oscillating_fit)
    np.testing.assert_allclose(mses, rgrid, rtol=RTOL)
    with pytest.raises(m.IHTBError):
        m.cv_run(y, g, z, folds, 3, [p + 1], d="Poisson", l="LogLink")


def test_properties_at_scale():
    """Size-independent checks at a size the oracle cannot hold: exactly k non-zeros, intercept estimated,
    FAST and EXACT sweeps give the same support / iterations and the same beta to 1e-9."""
    n, p, k = 20000, 60000, 10
    g = m.B200SnpLinAlg.synthetic(n, p, 77)
    y, z, idx, beta, c = synth.simulate_response(78, n, p, k, "Bernoulli", geno_seed=77)
    a = m.fit_iht(y, g, z, k=k, d="Bernoulli", l="LogitLink", sweep_mode=m.SWEEP_FAST)
    b = m.fit_iht(y, g, z, k=k, d="Bernoulli", l="LogitLink", sweep_mode=m.SWEEP_EXACT)
    assert np.count_nonzero(a.beta) == k and a.c[0] != 0
    assert a.iter == b.iter and np.array_equal(np.flatnonzero(a.beta), np.flatnonzero(b.beta))
    np.testing.assert_allclose(a.beta, b.beta, rtol=1e-9, atol=1e-12)
    assert abs(a.logl - b.logl) < 1e-9 * abs(b.logl)
    strong = idx[np.abs(beta) > 0.25]
    assert set(strong) <= set(np.flatnonzero(a.beta))


def test_dimension_errors():
    g = m.B200SnpLinAlg.synthetic(100, 50, 1)
    with pytest.raises(m.DimensionMismatch):
        m.fit_iht(np.zeros(99), g, None, k=3)
    with pytest.raises(m.DimensionMismatch):
        m.fit_iht(np.zeros(100), g, np.ones((100, 2)), k=3, zkeep=np.array([True]))
    with pytest.raises(AssertionError):
        m.fit_iht(np.zeros(100), g, None, k=-1)
    with pytest.raises(ValueError):
        m.cv_iht(np.zeros(100), g, None, path=[60], folds=np.ones(100, int))


def test_file_wrappers_on_bundled_plink(tmp_path, normal_data):
    """iht("normal", 7, Normal, covariates="covariates.txt") through the file-level wrapper == docs trace."""
    import shutil
    gold = json.load(open(os.path.join(GOLDEN, "docs_trace_normal_k7.json")))
    shutil.copy(os.path.join(GOLDEN, "normal.bed"), tmp_path / "normal.bed")
    with open(tmp_path / "normal.fam", "w") as f:
        for i, yi in enumerate(normal_data["y"]):
            f.write(f"{i + 1}\t1\t0\t0\t1\t{float(yi)!r}\n")
    shutil.copy(os.path.join(GOLDEN, "covariates.txt"), tmp_path / "covariates.txt")
    res = m.iht(str(tmp_path / "normal"), 7, "Normal", covariates=str(tmp_path / "covariates.txt"),
                summaryfile=str(tmp_path / "iht.summary.txt"), betafile=str(tmp_path / "iht.beta.txt"))
    assert res.iter == gold["iter"] and list(np.flatnonzero(res.beta) + 1) == gold["support_1based"]
    np.testing.assert_allclose([t[0] for t in res.trace], gold["logl"], rtol=1e-9)
    assert os.path.exists(tmp_path / "iht.summary.txt") and np.loadtxt(tmp_path / "iht.beta.txt").shape == (10000,)
    folds = synth.folds_for(1, 1000, 3)
    mse = m.cross_validate(str(tmp_path / "normal"), "Normal", path=[5, 7, 9], q=3,
                           covariates=str(tmp_path / "covariates.txt"), folds=folds)
    assert mse.shape == (3,) and np.all(mse > 0)


# ---- multivariate Normal (BASELINE config 4 family) ---------------------------------------------------------------
def _mv_data(seed, n, p, r, k):
    rng = np.random.default_rng(seed)
    bed = synth.packed_columns(seed, n, np.arange(p))
    idx = np.sort(rng.permutation(p)[:k])
    xs = synth.standardized_columns(seed, n, idx)
    B = np.zeros((r, k))
    for c in range(k):
        B[rng.integers(0, r), c] = rng.normal() * 0.8
    A = rng.normal(size=(r, r)); cov = A @ A.T / r + np.eye(r) * 0.5
    E = np.linalg.cholesky(cov) @ rng.normal(size=(r, n))
    Z = np.vstack([np.ones(n), rng.normal(size=n)])
    Cm = rng.normal(size=(r, 2)) * 0.5
    Y = B @ xs.T + Cm @ Z + E
    return bed, Y, Z


@pytest.mark.parametrize("mode", MODES + [m.SWEEP_PAIR])
@pytest.mark.parametrize("n,p,r,k", [(1200, 2500, 2, 8), (1500, 2000, 3, 9), (1003, 1500, 5, 10), (900, 1200, 18, 12)])
def test_mv_fit_matches_oracle(n, p, r, k, mode):
    from oracle import mviht
    bed, Y, Z = _mv_data(40 + r, n, p, r, k)
    g = m.B200SnpLinAlg.from_bed_columns(bed, n)
    res = m.fit_iht(Y, g, Z, k=k + 2, sweep_mode=mode)
    ref = mviht.fit_mv_iht(Y, snp.SnpLinAlgOracle(bed, n), Z, k=k + 2)
    assert isinstance(res, m.mIHTResult) and res.traits == r
    _compare_mv(res, ref)


def test_mv_init_beta_matches_oracle():
    """init_beta = true for MvNormal (src/multivariate.jl:425-429, 519-558), also under a CV mask."""
    from oracle import mviht
    n, p, r, k = 1200, 2000, 3, 9
    bed, Y, Z = _mv_data(77, n, p, r, k)
    g = m.B200SnpLinAlg.from_bed_columns(bed, n)
    o = snp.SnpLinAlgOracle(bed, n)
    for mode in MODES:
        res = m.fit_iht(Y, g, Z, k=k + 1, init_beta=True, sweep_mode=mode)
        ref = mviht.fit_mv_iht(Y, o, Z, k=k + 1, init_beta=True)
        _compare_mv(res, ref)
    plain = m.fit_iht(Y, g, Z, k=k + 1)
    assert res.iter != plain.iter or not np.array_equal(res.beta, plain.beta)
    folds = synth.folds_for(3, n, 2)
    mses, iters = m.cv_iht(Y, g, Z, path=[4, 9], q=2, folds=folds, init_beta=True, return_grid=True)
    _, rgrid, riters = ocv.cv_iht(Y, o, Z, path=[4, 9], q=2, folds=folds, init_beta=True, return_grid=True)
    assert np.array_equal(iters, riters)
    np.testing.assert_allclose(mses, rgrid, rtol=RTOL)


def test_mv_bundled_fixture():
    """Bundled data/multivariate.* (r = 2), k = 10: same support as the survey probe / oracle."""
    from oracle import mviht
    Y = np.loadtxt(os.path.join(GOLDEN, "multivariate_y.txt")).T
    n = Y.shape[1]
    bed = snp.read_bed(os.path.join(GOLDEN, "multivariate.bed"), n)
    g = m.B200SnpLinAlg.from_bed_columns(bed, n)
    res = m.fit_iht(Y, g, None, k=10)
    ref = mviht.fit_mv_iht(Y, snp.SnpLinAlgOracle(bed, n), None, k=10)
    _compare_mv(res, ref)
    assert list(np.flatnonzero(res.beta[0]) + 1) == [134, 442, 450, 1891, 2557, 3243, 3931, 9289]
    assert list(np.flatnonzero(res.beta[1]) + 1) == [1014, 5214]


def test_mv_cv_and_errors():
    from oracle import cv as ocv2
    bed, Y, Z = _mv_data(77, 900, 1200, 2, 5)
    g = m.B200SnpLinAlg.from_bed_columns(bed, 900)
    folds = synth.folds_for(3, 900, 3)
    mses, iters = m.cv_iht(Y, g, Z, path=[2, 5, 8], q=3, folds=folds, return_grid=True)
    _, rgrid, riters = ocv2.cv_iht(Y, snp.SnpLinAlgOracle(bed, 900), Z, path=[2, 5, 8], q=3, folds=folds,
                                   return_grid=True)
    assert np.array_equal(iters, riters)
    np.testing.assert_allclose(mses, rgrid, rtol=RTOL)
    with pytest.raises(m.DimensionMismatch):
        m.fit_iht(Y[:, :-1], g, Z, k=3)          # test/multivariate_test.jl:109-110


@pytest.mark.parametrize("est_r", ["MM", "Newton"])
def test_negbin_nuisance_estimation(est_r):
    """test/L0_reg_test.jl:245-297: est_r = :MM / :Newton for NegativeBinomial; compared with the oracle's mle_for_r."""
    n, p, k = 1500, 2000, 5
    y, z, _, _, _ = synth.simulate_response(91, n, p, k, "NegativeBinomial", nb_r=5.0)
    bed = synth.packed_columns(91, n, np.arange(p))
    g = m.B200SnpLinAlg.from_bed_columns(bed, n)
    res = m.fit_iht(y, g, z, k=k + 1, d="NegativeBinomial", l="LogLink", est_r=est_r, nb_r=1.0)
    ref = iht.fit_iht(y, snp.SnpLinAlgOracle(bed, n), z, k=k + 1, d=glm.NEGBIN, l=glm.LOG, est_r=est_r, nb_r=1.0)
    assert res.iter == ref.iter and np.array_equal(np.flatnonzero(res.beta), np.flatnonzero(ref.beta))
    np.testing.assert_allclose(res.beta, ref.beta, rtol=1e-5, atol=1e-10)
    assert abs(res.logl - ref.logl) <= 1e-6 * abs(ref.logl)
    with pytest.raises(m.IHTBError):
        m.fit_iht(y, g, z, k=3, d="Poisson", l="LogLink", est_r="MM")


@pytest.mark.parametrize("d,l", [("Bernoulli", "ProbitLink"), ("Bernoulli", "CloglogLink"), ("Normal", "LogLink")])
def test_non_canonical_links(d, l):
    """docs/src/index.md "Available link functions": the device GLM kernels carry all nine GLM.jl links."""
    n, p, k = 2000 if d == "Bernoulli" else 1500, 2000, 4
    seed = 500 + n
    y, z, *_ = synth.simulate_response(seed, n, p, k, d, n_cov=1)
    if d == "Normal":
        y = np.abs(y) + 0.5
    bed = synth.packed_columns(seed, n, np.arange(p))
    g = m.B200SnpLinAlg.from_bed_columns(bed, n)
    kfit = 4 if d == "Bernoulli" else 5
    res = m.fit_iht(y, g, z, k=kfit, d=d, l=l)
    ref = iht.fit_iht(y, snp.SnpLinAlgOracle(bed, n), z, k=kfit, d=d, l=l)
    assert ref.iter < 200
    _compare(res, ref, rtol=1e-5)


def test_edge_cases():
    """k = p, k = 1, a monomorphic column (sigma_inv = 1), a column that is entirely missing is not supported by the
    reference either (mu = NaN) so it is not exercised; p < 128 (one partial column block), n < 512 (one partial slab)."""
    rng = np.random.default_rng(0)
    n, p = 301, 37
    g_ = rng.integers(0, 3, size=(n, p)); g_[:, 5] = 0                      # monomorphic SNP
    bed = snp.pack_codes(snp.dosage_to_codes(g_))
    o = snp.SnpLinAlgOracle(bed, n)
    x = m.B200SnpLinAlg.from_bed_columns(bed, n)
    mu, sinv, _ = x.stats()
    assert sinv[5] == 1.0 and mu[5] == 0.0
    y = o.dense()[:, [3, 20]] @ np.array([1.0, -0.7]) + rng.normal(size=n) * 0.5 + 2.0
    for k in (1, 2, 10, p):
        res = m.fit_iht(y, x, None, k=k)
        ref = iht.fit_iht(y, o, None, k=k)
        # with k = 1 the iterate sits on a fixed point after two steps: old and new loglikelihood are equal to the
        # last ulp, so `prev_logl > logl` (src/utilities.jl:485) is decided by rounding noise -> backtrack counts
        # are not comparable there (model, loglikelihood and iteration count still are)
        _compare(res, ref, check_backtracks=(k != 1))
    v = rng.normal(size=n)
    np.testing.assert_allclose(x.xt_v(v, m.SWEEP_FAST), o.xt_v(v), atol=1e-4 * np.abs(o.xt_v(v)).max())
    np.testing.assert_allclose(x.xt_v(v, m.SWEEP_EXACT), o.xt_v(v), atol=1e-11 * np.abs(o.xt_v(v)).max())


def test_init_beta_matches_oracle():
    """test/L0_reg_test.jl:299-321 (init_beta = true): beta starts from univariate regressions (initialize_beta!)."""
    n, p, k = 1200, 2500, 6
    y, z, _, _, _ = synth.simulate_response(61, n, p, k, "Normal", n_cov=2, missing_rate=0.01)
    bed = synth.packed_columns(61, n, np.arange(p), 0.01)
    g = m.B200SnpLinAlg.from_bed_columns(bed, n)
    o = snp.SnpLinAlgOracle(bed, n)
    for mode in MODES:
        res = m.fit_iht(y, g, z, k=k + 1, init_beta=True, sweep_mode=mode)
        ref = iht.fit_iht(y, o, z, k=k + 1, init_beta=True)
        _compare(res, ref)
    plain = m.fit_iht(y, g, z, k=k + 1)
    assert res.iter != plain.iter or not np.array_equal(res.beta, plain.beta)      # it really is a different start
    # under a CV mask, and through cv_iht
    folds = synth.folds_for(9, n, 3)
    mses, iters = m.cv_iht(y, g, z, path=[3, 7], q=3, folds=folds, init_beta=True, return_grid=True)
    _, rgrid, riters = ocv.cv_iht(y, o, z, path=[3, 7], q=3, folds=folds, init_beta=True, return_grid=True)
    assert np.array_equal(iters, riters)
    np.testing.assert_allclose(mses, rgrid, rtol=RTOL)
    with pytest.raises(m.IHTBError):
        m.fit_iht((y > 0).astype(float), g, z, k=3, d="Bernoulli", l="LogitLink", init_beta=True)


WEIGHT_CASES = [
    ("Normal", "IdentityLink", 1200, 3000, 8, 10, 2, 0.0),
    ("Bernoulli", "LogitLink", 2000, 3000, 5, 7, 1, 0.0),
    ("Poisson", "LogLink", 1500, 2500, 6, 8, 1, 0.005),
]


@pytest.mark.parametrize("d,l,n,p,k,kfit,ncov,miss", WEIGHT_CASES)
def test_prior_weights_match_oracle(d, l, n, p, k, kfit, ncov, miss):
    """`weight` keyword (src/fit.jl:69, docs 'maf_weights'): the projection ranks |b_j| * w_j (src/utilities.jl:291-354)."""
    seed = 100 + n + p
    y, z, _, _, _ = synth.simulate_response(seed, n, p, k, d, n_cov=ncov, missing_rate=miss)
    bed = synth.packed_columns(seed, n, np.arange(p), miss)
    g = m.B200SnpLinAlg.from_bed_columns(bed, n)
    o = snp.SnpLinAlgOracle(bed, n)
    w = m.maf_weights(g, max_weight=5.0)
    maf = np.minimum(o.mu / 2, 1 - o.mu / 2)
    np.testing.assert_array_equal(g.maf(), maf)                       # SnpArrays maf(), bit for bit
    np.testing.assert_array_equal(w, np.clip(1 / (2 * np.sqrt(maf * (1 - maf))), 1.0, 5.0))
    plain = m.fit_iht(y, g, z, k=kfit, d=d, l=l)
    for mode in MODES:
        res = m.fit_iht(y, g, z, k=kfit, d=d, l=l, weight=w, sweep_mode=mode)
        ref = iht.fit_iht(y, o, z, k=kfit, d=d, l=l, weight=w)
        assert ref.iter < 200
        _compare(res, ref)
    assert not np.array_equal(np.flatnonzero(res.beta), np.flatnonzero(plain.beta))    # the weights do matter here
    # a fit without weights after a weighted one on the same cached workspace is unweighted again
    again = m.fit_iht(y, g, z, k=kfit, d=d, l=l)
    np.testing.assert_array_equal(again.beta, plain.beta)
    with pytest.raises(m.DimensionMismatch):
        m.fit_iht(y, g, z, k=kfit, d=d, l=l, weight=w[:-1])
    with pytest.raises(m.IHTBError):
        m.fit_iht(y, g, z, k=kfit, d=d, l=l, weight=-w)


def test_prior_weights_cv_and_init_beta():
    n, p, k = 1200, 2000, 5
    y, z, _, _, _ = synth.simulate_response(77, n, p, k, "Normal", n_cov=1)
    bed = synth.packed_columns(77, n, np.arange(p))
    g = m.B200SnpLinAlg.from_bed_columns(bed, n)
    o = snp.SnpLinAlgOracle(bed, n)
    w = m.maf_weights(g, max_weight=4.0)
    folds = synth.folds_for(5, n, 3)
    mses, iters = m.cv_iht(y, g, z, path=[2, 5, 8], q=3, folds=folds, weight=w, return_grid=True)
    _, rgrid, riters = ocv.cv_iht(y, o, z, path=[2, 5, 8], q=3, folds=folds, weight=w, return_grid=True)
    assert np.array_equal(iters, riters)
    np.testing.assert_allclose(mses, rgrid, rtol=RTOL)
    res = m.fit_iht(y, g, z, k=6, weight=w, init_beta=True)
    ref = iht.fit_iht(y, o, z, k=6, weight=w, init_beta=True)
    _compare(res, ref)


DEBIAS_CASES = [
    ("Normal", "IdentityLink", 1200, 3000, 8, 10, 2, 0.0, {}),
    ("Bernoulli", "LogitLink", 2000, 3000, 5, 7, 1, 0.0, {}),
    ("Bernoulli", "ProbitLink", 2000, 2500, 5, 6, 1, 0.005, {"max_iter": 25}),
    ("Poisson", "LogLink", 1500, 2500, 6, 8, 1, 0.0, {"max_iter": 25}),
    ("NegativeBinomial", "LogLink", 1500, 2500, 6, 8, 0, 0.0, {"max_iter": 12}),
]


@pytest.mark.parametrize("d,l,n,p,k,kfit,ncov,miss,kw", DEBIAS_CASES)
def test_debias_matches_oracle(d, l, n, p, k, kfit, ncov, miss, kw):
    """`debias=true` (src/fit.jl:187-188, src/utilities.jl:1014-1020): when the support did not change (iteration >= 5)
    beta[idx] is replaced by GLM.jl's IRLS fit of y on x[:, idx].  GLM.jl is not in the reference tree: the oracle
    restates its `_fit!` loop (oracle/glm.py::glm_fit), so this parity is oracle-pinned only."""
    seed = 100 + n + p
    y, z, _, _, _ = synth.simulate_response(seed, n, p, k, d, n_cov=ncov, missing_rate=miss)
    bed = synth.packed_columns(seed, n, np.arange(p), miss)
    g = m.B200SnpLinAlg.from_bed_columns(bed, n)
    o = snp.SnpLinAlgOracle(bed, n)
    plain = m.fit_iht(y, g, z, k=kfit, d=d, l=l, nb_r=10.0, **kw)
    for mode in MODES:
        res = m.fit_iht(y, g, z, k=kfit, d=d, l=l, nb_r=10.0, debias=True, sweep_mode=mode, **kw)
        ref = iht.fit_iht(y, o, z, k=kfit, d=d, l=l, nb_r=10.0, debias=True, **kw)
        _compare(res, ref)
    assert not np.array_equal(res.beta, plain.beta)


def test_debias_cv_and_mv_error():
    n, p, k = 1500, 2000, 5
    y, z, _, _, _ = synth.simulate_response(31, n, p, k, "Bernoulli", n_cov=0)
    bed = synth.packed_columns(31, n, np.arange(p))
    g = m.B200SnpLinAlg.from_bed_columns(bed, n)
    o = snp.SnpLinAlgOracle(bed, n)
    folds = synth.folds_for(3, n, 3)
    # test/cv_iht_test.jl:26-29: cv_iht(..., debias=true, max_iter=10)
    mses, iters = m.cv_iht(y, g, z, d="Bernoulli", l="LogitLink", path=[3, 6], q=3, folds=folds, debias=True,
                           max_iter=10, return_grid=True)
    _, rgrid, riters = ocv.cv_iht(y, o, z, d="Bernoulli", l="LogitLink", path=[3, 6], q=3, folds=folds, debias=True,
                                  max_iter=10, return_grid=True)
    assert np.array_equal(iters, riters) and np.all(mses > 0)
    np.testing.assert_allclose(mses, rgrid, rtol=RTOL)
    Y = np.vstack([y, 1 - y])
    with pytest.raises(m.IHTBError, match="debiasing routine for multivariate"):
        m.fit_iht(Y, g, np.ones((1, n)), k=4, debias=True)


GROUP_CASES = [("Normal", "IdentityLink"), ("Bernoulli", "LogitLink"), ("Poisson", "LogLink")]


@pytest.mark.parametrize("d,l", GROUP_CASES)
def test_group_projection_matches_oracle(d, l):
    """Doubly sparse projection (J groups x k predictors, src/utilities.jl:613-679; test/L0_reg_test.jl:185-243):
    contiguous blocks, per-group k vector (empty initial support, :426-430) and scattered membership."""
    n, p, k = 1500, 3000, 6
    y, z, _, _, _ = synth.simulate_response(41, n, p, k, d, n_cov=1)
    bed = synth.packed_columns(41, n, np.arange(p))
    g = m.B200SnpLinAlg.from_bed_columns(bed, n)
    o = snp.SnpLinAlgOracle(bed, n)
    blocks = np.arange(p) // 100 + 1
    ks = [2] * 30; ks[3] = 1
    scattered = np.random.default_rng(0).integers(1, 8, p)
    for kw in ({"k": 2, "J": 4, "group": blocks}, {"k": ks, "J": 3, "group": blocks}, {"k": 2, "J": 3, "group": scattered},
               {"k": 3, "J": 1, "group": np.ones(p, dtype=int)}):
        res = m.fit_iht(y, g, z, d=d, l=l, **kw)
        ref = iht.fit_iht(y, o, z, d=d, l=l, **kw)
        assert ref.iter < 200
        _compare(res, ref)
        nz = np.flatnonzero(res.beta)
        grp = np.asarray(kw["group"])
        assert len(set(grp[nz])) <= kw["J"]
        kmax = kw["k"] if np.isscalar(kw["k"]) else None
        for gid in set(grp[nz]):
            assert (grp[nz] == gid).sum() <= (kmax if kmax is not None else kw["k"][gid - 1])
    # one group that holds everything == plain top-k projection after the first iteration's start
    assert np.count_nonzero(res.beta) <= 3


def test_group_projection_cv_weights_errors():
    n, p, k = 1200, 2000, 5
    y, z, _, _, _ = synth.simulate_response(43, n, p, k, "Normal", n_cov=1)
    bed = synth.packed_columns(43, n, np.arange(p))
    g = m.B200SnpLinAlg.from_bed_columns(bed, n)
    o = snp.SnpLinAlgOracle(bed, n)
    blocks = np.arange(p) // 50 + 1
    folds = synth.folds_for(4, n, 3)
    mses, iters = m.cv_iht(y, g, z, path=[1, 2, 3], q=3, folds=folds, J=3, group=blocks, return_grid=True)
    _, rgrid, riters = ocv.cv_iht(y, o, z, path=[1, 2, 3], q=3, folds=folds, J=3, group=blocks, return_grid=True)
    assert np.array_equal(iters, riters)
    np.testing.assert_allclose(mses, rgrid, rtol=RTOL)
    w = m.maf_weights(g, max_weight=3.0)           # weights only shape the initial (ungrouped) support in group mode
    _compare(m.fit_iht(y, g, z, k=2, J=3, group=blocks, weight=w, debias=True),
             iht.fit_iht(y, o, z, k=2, J=3, group=blocks, weight=w, debias=True))
    with pytest.raises(m.DimensionMismatch):
        m.fit_iht(y, g, z, k=2, J=3, group=blocks[:-1])
    with pytest.raises(AssertionError):
        m.fit_iht(y, g, z, k=[1, 2])
    with pytest.raises(m.IHTBError):                # check_group: a group with no more members than its k
        m.fit_iht(y, g, z, k=[50] * 40, J=2, group=blocks)
    with pytest.raises(AssertionError):
        m.fit_iht(y, g, z, k=2, J=-1, group=blocks)


def test_next_tier_options_against_frozen_fixtures(normal_data):
    """CUDA path vs tests/golden/oracle_regression.json (weights, debias, groups, per-group k, init_beta on the bundled
    1000 x 10000 PLINK file): support and iteration counts identical, values within 1e-6."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("mk", os.path.join(GOLDEN, "make_oracle_regression.py"))
    mk = importlib.util.module_from_spec(spec); spec.loader.exec_module(mk)
    frozen = json.load(open(os.path.join(GOLDEN, "oracle_regression.json")))
    g = m.B200SnpLinAlg.from_bed_file(os.path.join(GOLDEN, "normal.bed"), 1000)
    for name, kw in mk.cases(g.p).items():
        res = m.fit_iht(normal_data["y"], g, None, **kw)
        want = frozen[name]
        nz = np.flatnonzero(res.beta)
        assert res.iter == want["iter"] and list(nz) == want["support_0based"], name
        np.testing.assert_allclose(res.beta[nz], want["beta"], rtol=RTOL, err_msg=name)
        np.testing.assert_allclose(res.c, want["c"], rtol=RTOL, err_msg=name)
        assert abs(res.logl - want["logl"]) <= RTOL * abs(want["logl"]), name
        assert abs(res.sigma_g - want["sigma_g"]) <= RTOL * abs(want["sigma_g"]), name


def test_genuine_ties_at_the_kth_magnitude():
    """Two IDENTICAL SNP columns have identical gradients and coefficients, so the k-th magnitude is tied whenever only
    one of them fits into the support.  The reference prunes tied entries at random (`_choose!`, src/utilities.jl:444-458);
    the documented deterministic rule (lowest position wins; oracle.iht.prune_ties) must hold on the device for the
    initial support, every gradient step and every backtrack -- same support, iterations and values as the oracle."""
    n, p, k = 900, 700, 5
    bed = synth.packed_columns(77, n, np.arange(p))
    xs = synth.standardized_columns(77, n, np.array([10, 50, 90, 400]))
    rng = np.random.default_rng(3)
    y = xs @ np.array([1.2, -0.9, 0.7, 0.5]) + rng.normal(size=n)
    for dup_src, dup_dst in [(10, 300), (400, 401), (90, 5)]:       # duplicate of a causal column after / right after / before it
        b2 = bed.copy()
        b2[dup_dst] = b2[dup_src]
        g = m.B200SnpLinAlg.from_bed_columns(b2, n)
        o = snp.SnpLinAlgOracle(b2, n)
        for kk in (1, 3, 4, 5):
            for mode in MODES:
                res = m.fit_iht(y, g, None, k=kk, sweep_mode=mode)
                ref = iht.fit_iht(y, o, None, k=kk)
                # two identical columns make the model degenerate: near convergence successive loglikelihoods agree to
                # the last bits, where `prev_logl > logl` (the backtracking test) is decided by summation order.
                # Backtrack counts are therefore compared on the iterations whose loglikelihood moves by more than
                # rounding; support, iterations and every value are compared everywhere.
                _compare(res, ref, check_backtracks=False)
                lg = np.asarray(ref.trace.logl)
                moved = np.r_[True, np.abs(np.diff(lg)) > 1e-10 * np.abs(lg[1:])]
                assert [t[1] for t, mv in zip(res.trace, moved) if mv] == [b for b, mv in zip(ref.trace.backtracks, moved) if mv]
                # with the intercept kept, a two-way tie is within k + zkeepn and is NOT pruned (src/utilities.jl:448-449)
                if kk == 1 and dup_src == 10:
                    assert res.beta[dup_src] != 0 and res.beta[dup_dst] != 0 and res.beta[dup_src] == res.beta[dup_dst]
        g.close()


def test_device_side_backtracking_choice_matches_default():
    """IHTB_FUSE=1: the backtracking winner is chosen on the device and the whole iteration needs one host round trip
    (fit.cu one_step_fused).  Same support, iterations, backtracks and values as the default two-round-trip step."""
    for d, l, n, p, k, ncov in [("Bernoulli", "LogitLink", 2000, 3000, 7, 1), ("Normal", "IdentityLink", 1200, 3000, 10, 2),
                                ("Poisson", "LogLink", 1500, 2500, 8, 1)]:
        seed = 100 + n + p
        y, z, *_ = synth.simulate_response(seed, n, p, k - 2, d, n_cov=ncov)
        g = m.B200SnpLinAlg.from_bed_columns(synth.packed_columns(seed, n, np.arange(p)), n)
        ref = m.fit_iht(y, g, z, k=k, d=d, l=l)
        os.environ["IHTB_FUSE"] = "1"
        try:
            res = m.fit_iht(y, g, z, k=k, d=d, l=l)
        finally:
            os.environ.pop("IHTB_FUSE", None)
        assert res.iter == ref.iter and [t[1] for t in res.trace] == [t[1] for t in ref.trace]
        assert np.array_equal(np.flatnonzero(res.beta), np.flatnonzero(ref.beta))
        np.testing.assert_allclose(res.beta, ref.beta, rtol=1e-12, atol=0)
        np.testing.assert_allclose([t[0] for t in res.trace], [t[0] for t in ref.trace], rtol=1e-12)
        assert res.n_launches != ref.n_launches            # it really took the other path
        g.close()


def test_cv_grid_does_not_depend_on_sweep_pairing():
    """ihtb_cv_run runs two fits at a time per device and serves their sweeps with one PAIR pass (half2 tables, looser
    error bound, more candidates re-scored exactly).  The grid must not depend on it -- plain, with prior weights and
    debiasing, and with an odd number of fits (the last one sweeps alone): same iteration counts, losses equal to 1e-10
    (every candidate is re-scored exactly by the same nibble-table kernel whatever sweep screened it)."""
    n, p, q = 3000, 6000, 3
    y, z, *_ = synth.simulate_response(31, n, p, 6, "Normal", n_cov=1)
    g = m.B200SnpLinAlg.synthetic(n, p, 31)
    folds = synth.folds_for(31, n, q)
    w = m.maf_weights(g, max_weight=3.0)
    for path, kw in (([1, 2, 3, 4, 5], {}), ([2, 4, 6], {"weight": w}), ([3, 5, 6], {"debias": True})):
        got = m.cv_run(y, g, z, folds, q, path, **kw)
        os.environ["IHTB_CV_PAIR"] = "0"
        try:
            want = m.cv_run(y, g, z, folds, q, path, **kw)
        finally:
            os.environ.pop("IHTB_CV_PAIR", None)
        assert np.array_equal(got[1], want[1]) and got[1].max() < 100          # converging fits (cf. test_oscillating_fit)
        np.testing.assert_allclose(got[0], want[0], rtol=1e-10, atol=0)
    g.close()


def test_long_vectors_adaptive_batch_matches_oracle():
    """n >= 131072: the gradient step and its backtracks are evaluated in a first batch sized from the previous step's
    backtracks, the remaining candidate models in a second round trip only when the walk needs them (fit.cu
    finish_batched); candidates are re-scored by the nibble-table gather.  The oracle's fit on this problem takes
    0, 1, 2 and 3 backtracks in turn, so every batch size and the second round trip are exercised."""
    from oracle import cpu as ocpu
    d, l, n, p, k, seed = "Bernoulli", "LogitLink", 140_000, 600, 5, 4243
    bed = ocpu.synth_columns(seed, n, 0, p)
    y, z, *_ = synth.simulate_response(seed + 1, n, p, k, d, geno_seed=seed)
    g = m.B200SnpLinAlg.from_bed_columns(bed, n)
    res = m.fit_iht(y, g, z, k=k + 1, d=d, l=l)
    ref = iht.fit_iht(y, ocpu.PackedSnpLinAlgCPU(bed, n), z, k=k + 1, d=d, l=l)
    assert ref.iter < 200 and set(ref.trace.backtracks) == {0, 1, 2, 3}
    _compare(res, ref)
    g.close()
