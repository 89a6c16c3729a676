"""Pins the CPU oracle to the reference's own known-answer vectors and to the closed forms
the reference's unit tests use.  CPU only."""
import json
import os

import numpy as np
import pytest

from oracle import cv as ocv
from oracle import glm, iht, mviht, snp
from conftest import GOLDEN


def test_docs_trace_normal_k7(normal_data, normal_oracle):
    """docs/src/man/examples.md:230-268: 5 iterations, logl/tol per iteration, support, beta, c, PVE."""
    gold = json.load(open(os.path.join(GOLDEN, "docs_trace_normal_k7.json")))
    res = iht.fit_iht(normal_data["y"], normal_oracle, normal_data["z"], k=7, d=glm.NORMAL, l=glm.IDENTITY)
    assert res.iter == gold["iter"]
    np.testing.assert_allclose(res.trace.logl, gold["logl"], rtol=0, atol=1e-10)
    np.testing.assert_allclose(res.trace.tol, gold["tol"], rtol=1e-12)
    assert res.trace.backtracks == gold["backtracks"]
    assert abs(res.logl - gold["final_logl"]) < 1e-10
    assert abs(res.sigma_g - gold["pve"]) < 1e-13
    nz = np.flatnonzero(res.beta)
    assert list(nz + 1) == gold["support_1based"]
    np.testing.assert_allclose(res.beta[nz], gold["beta_6sig"], rtol=2e-6)
    np.testing.assert_allclose(res.c, gold["c_6sig"], rtol=2e-6)


def test_stale_k8_output_support(normal_data, normal_oracle):
    """data/iht.beta.txt (older release) selected SNPs 3136,3137,4246,4717,6290,7755,8375,9415; the current
    algorithm with k=8 finds the same support and optimum."""
    res = iht.fit_iht(normal_data["y"], normal_oracle, normal_data["z"], k=8)
    assert list(np.flatnonzero(res.beta) + 1) == [3136, 3137, 4246, 4717, 6290, 7755, 8375, 9415]
    assert abs(res.logl - (-1390.3003586)) < 1e-5


def test_true_beta_recovered(normal_data, normal_oracle):
    """data/normal_true_beta.txt holds the simulation truth: the k=9 README fit recovers large effects."""
    rows = [l.strip().split(",") for l in open(os.path.join(GOLDEN, "normal_true_beta.txt"))][1:]
    truth = {int(r[0][3:]) - 1: float(r[1]) for r in rows}
    res = iht.fit_iht(normal_data["y"], normal_oracle, None, k=9)
    nz = np.flatnonzero(res.beta)
    big = {j for j, b in truth.items() if abs(b) > 0.3}
    assert big <= set(nz)
    for j in big:
        assert abs(res.beta[j] - truth[j]) < 0.1
    assert res.iter == 10 and abs(res.logl - (-1612.734968053745)) < 1e-8


def test_bundled_bed_has_no_missing(normal_data):
    codes = snp.unpack_codes(normal_data["bed"], normal_data["n"])
    assert (codes == 1).sum() == 0
    assert [(codes == c).sum() for c in (0, 2, 3)] == [5807184, 3352153, 840663]


def test_dense_standardization_equals_snplinalg(normal_data, normal_oracle):
    """test/wrapper_test.jl:186-194: convert(Matrix, SnpArray; center, scale, impute) == standardize_genotypes!."""
    g = snp.dosages(normal_data["bed"], normal_data["n"])
    mu = np.nanmean(g, axis=0)
    sd = np.sqrt(mu * (1 - mu / 2))
    dense = (np.where(np.isnan(g), mu, g) - mu) / np.where(sd > 0, sd, 1)
    np.testing.assert_allclose(normal_oracle.dense(), dense, rtol=1e-13, atol=1e-13)
    v = np.random.default_rng(0).normal(size=normal_data["n"])
    np.testing.assert_allclose(normal_oracle.xt_v(v), dense.T @ v, rtol=1e-9, atol=1e-9)


def test_pack_unpack_roundtrip_with_missing():
    rng = np.random.default_rng(1)
    g = rng.integers(-1, 3, size=(1003, 17))
    bed = snp.pack_codes(snp.dosage_to_codes(g))
    d = snp.dosages(bed, 1003)
    assert np.array_equal(np.isnan(d), g < 0)
    assert np.array_equal(d[g >= 0], g[g >= 0].astype(float))
    mu, sinv, nm = snp.column_stats(bed, 1003)
    assert np.array_equal(nm, (g < 0).sum(axis=0))
    np.testing.assert_allclose(mu, np.where(g < 0, 0, g).sum(axis=0) / (1003 - nm))


# ---- closed forms the reference's unit tests assert (test/utilities_test.jl:20-92) -----------------
@pytest.mark.parametrize("d,l", [(glm.NORMAL, glm.IDENTITY), (glm.BERNOULLI, glm.LOGIT), (glm.POISSON, glm.LOG),
                                 (glm.NEGBIN, glm.LOG)])
def test_loglikelihood_is_sum_of_logpdf(d, l):
    from scipy import stats
    rng = np.random.default_rng(3)
    n = 500
    eta = rng.normal(scale=0.5, size=n)
    mu = glm.linkinv(l, eta)
    if d == glm.NORMAL:
        y = mu + rng.normal(size=n)
        phi = np.sum((y - mu) ** 2) / n
        ref = stats.norm.logpdf(y, mu, np.sqrt(phi)).sum()
    elif d == glm.BERNOULLI:
        y = (rng.random(n) < mu).astype(float)
        ref = stats.bernoulli.logpmf(y, mu).sum()
    elif d == glm.POISSON:
        y = rng.poisson(mu).astype(float)
        ref = stats.poisson.logpmf(y, mu).sum()
    else:
        r = 10.0
        y = rng.negative_binomial(r, r / (mu + r)).astype(float)
        ref = stats.nbinom.logpmf(y, r, r / (mu + r)).sum()
    w = np.ones(n)
    got = glm.loglikelihood(d, y, mu, w, r=10.0)
    assert abs(got - ref) < 1e-8
    if d == glm.NORMAL:
        assert abs(glm.deviance(d, y, mu, w) - np.sum((y - mu) ** 2)) < 1e-9


def test_update_mu_links():
    eta = np.linspace(-3, 3, 13)
    np.testing.assert_allclose(glm.linkinv(glm.LOGIT, eta), np.exp(eta) / (1 + np.exp(eta)), atol=1e-12)
    np.testing.assert_allclose(glm.linkinv(glm.LOG, eta), np.exp(eta), atol=1e-12)
    np.testing.assert_allclose(glm.linkinv(glm.IDENTITY, eta), eta)
    for l in (glm.LOGIT, glm.LOG, glm.PROBIT, glm.CLOGLOG, glm.CAUCHIT, glm.SQRT):
        h = 1e-6
        num = (glm.linkinv(l, eta + h) - glm.linkinv(l, eta - h)) / (2 * h)
        np.testing.assert_allclose(glm.mueta(l, eta), num, rtol=1e-6, atol=1e-8)


def test_project_k_keeps_topk():
    """test/utilities_test.jl:166-176."""
    x = np.random.default_rng(5).random(100000)
    ref = np.sort(x)[-100:]
    iht.project_k(x, 100)
    assert np.count_nonzero(x) == 100
    np.testing.assert_array_equal(np.sort(x[x != 0]), ref)


def test_backtrack_truth_table():
    """test/utilities_test.jl:133-141: (prev_logl > logl) && (eta_step < nstep)."""
    f = lambda logl, prev, step, nstep: (prev > logl) and (step < nstep)
    assert f(-10.0, -5.0, 0, 3) and not f(-5.0, -10.0, 0, 3) and not f(-10.0, -5.0, 3, 3)


def test_meanloss_and_grid():
    folds = np.array([1, 1, 2, 2, 2, 3])
    grid = ocv.allocate_fold_and_k(3, [1, 5])
    assert grid == [(1, 1), (1, 5), (2, 1), (2, 5), (3, 1), (3, 5)]
    loss = ocv.meanloss(np.array([1.0, 2, 3, 4, 5, 6]), 3, folds)
    np.testing.assert_allclose(loss, [1 * 2 / 6 + 3 * 3 / 6 + 5 / 6, 2 * 2 / 6 + 4 * 3 / 6 + 6 / 6])


def test_pivoted_cholesky_matches_lapack_semantics():
    rng = np.random.default_rng(7)
    a = rng.normal(size=(5, 5)); a = a @ a.T + np.diag([5, 1, 9, 2, 7.0])
    u = mviht.pivoted_cholesky_upper(a)
    assert np.allclose(np.tril(u, -1), 0)
    # U'U reproduces a symmetric permutation of A whose diagonal pivots are non-increasing
    g = u.T @ u
    assert np.allclose(np.sort(np.diag(g)), np.sort(np.diag(a)))
    assert np.allclose(np.sort(np.linalg.eigvalsh(g)), np.sort(np.linalg.eigvalsh(a)))
    assert abs(u[0, 0] ** 2 - a.diagonal().max()) < 1e-12
    # identical to plain Cholesky when the diagonal is already descending and dominant
    d = np.diag([9.0, 5, 3, 2, 1]) + 0.01
    assert np.allclose(mviht.pivoted_cholesky_upper(d), np.linalg.cholesky(d).T)


def test_multivariate_bundled_fixture():
    """SURVEY.md App. C probe: bundled multivariate.* with k=10 (oracle-derived; the reference docs run used a
    different phenotype file, so this only pins the oracle against regressions and the true covariance)."""
    Y = np.loadtxt(os.path.join(GOLDEN, "multivariate_y.txt")).T
    n = Y.shape[1]
    bed = snp.read_bed(os.path.join(GOLDEN, "multivariate.bed"), n)
    x = snp.SnpLinAlgOracle(bed, n)
    res = mviht.fit_mv_iht(Y, x, None, k=10)
    assert res.iter >= 5 and np.all(res.sigma_g > 0)
    assert np.count_nonzero(res.beta) == 10
    t1 = list(np.flatnonzero(res.beta[0]) + 1); t2 = list(np.flatnonzero(res.beta[1]) + 1)
    assert t1 == [134, 442, 450, 1891, 2557, 3243, 3931, 9289] and t2 == [1014, 5214]
    true_cov = np.array([[0.9555626544452652, -0.08844663504667233], [-0.08844663504667233, 1.6257269206489926]])
    assert np.max(np.abs(res.Sigma - true_cov)) < 0.15


def test_cv_small_all_positive(normal_data, normal_oracle):
    """test/cv_iht_test.jl: all(mses .> 0)."""
    folds = 1 + (np.arange(normal_data["n"]) % 3)
    mse = ocv.cv_iht(normal_data["y"], normal_oracle, normal_data["z"], path=[1, 5, 7], q=3, folds=folds)
    assert mse.shape == (3,) and np.all(mse > 0)
    assert np.argmin(mse) == 2


def test_initialize_beta_is_univariate_least_squares(normal_data, normal_oracle):
    """`initialize_beta!` / `linreg!` (src/utilities.jl:776-842): slope of y ~ 1 + x_j, clamped to [-2, 2]; intercept
    averaged over all p + q - 1 regressions."""
    y, z, n = normal_data["y"], normal_data["z"], normal_data["n"]
    v = iht.IHTVariable(normal_oracle, z, y, 7, glm.NORMAL, glm.IDENTITY)
    v.init_iht_indices(np.ones(n, bool))
    v.b[:] = 0
    v.initialize_beta(np.ones(n, bool))
    icpts = []
    for j in (0, 17, 4716, 9414):
        X = np.c_[np.ones(n), normal_oracle.dense()[:, j]]
        sol = np.linalg.lstsq(X, y, rcond=None)[0]
        assert abs(np.clip(sol[1], -2, 2) - v.b[j]) < 1e-12
    res = iht.fit_iht(y, normal_oracle, z, k=7, init_beta=True)
    assert list(np.flatnonzero(res.beta) + 1) == [3137, 4246, 4717, 6290, 7755, 8375, 9415]     # same optimum
    assert abs(res.logl - (-1397.88074)) < 1e-4


def test_glm_fit_restates_irls():
    """oracle/glm.py::glm_fit (GLM.jl `_fit!`, used by debias!): exact OLS for Normal/identity in one Newton step, a
    stationary point of the likelihood for the other families, `linkfun` inverts `linkinv`."""
    rng = np.random.default_rng(1)
    n, k = 800, 4
    X = rng.standard_normal((n, k)); bt = 0.4 * rng.standard_normal(k)
    y = X @ bt + rng.standard_normal(n)
    np.testing.assert_allclose(glm.glm_fit(X, y, glm.NORMAL, glm.IDENTITY), np.linalg.lstsq(X, y, rcond=None)[0], rtol=1e-10)
    for d, l in [(glm.BERNOULLI, glm.LOGIT), (glm.POISSON, glm.LOG), (glm.BERNOULLI, glm.PROBIT)]:
        mu = glm.linkinv(l, X @ bt)
        yy = (rng.random(n) < mu).astype(float) if d == glm.BERNOULLI else rng.poisson(mu).astype(float)
        b = glm.glm_fit(X, yy, d, l)
        eta = X @ b
        score = X.T @ ((yy - glm.linkinv(l, eta)) * glm.mueta(l, eta) / glm.glmvar(d, glm.linkinv(l, eta)))
        assert np.abs(score).max() < 0.05 * np.sqrt(n)              # stops on the deviance criterion, not the score
    for l in (glm.IDENTITY, glm.LOGIT, glm.LOG, glm.PROBIT, glm.CLOGLOG, glm.CAUCHIT, glm.SQRT, glm.INVERSE, glm.INVSQ):
        mu = np.array([0.2, 0.5, 0.9])
        np.testing.assert_allclose(glm.linkinv(l, glm.linkfun(l, mu)), mu, rtol=1e-12)


def test_debias_and_weights_on_bundled_data(normal_data, normal_oracle):
    """Debiasing (src/fit.jl:187-188) refits the support without shrinkage; weights rescale the projection only."""
    y = normal_data["y"]
    plain = iht.fit_iht(y, normal_oracle, None, k=9)
    deb = iht.fit_iht(y, normal_oracle, None, k=9, debias=True)
    assert np.array_equal(np.flatnonzero(deb.beta), np.flatnonzero(plain.beta))
    assert deb.iter <= plain.iter
    ones = iht.fit_iht(y, normal_oracle, None, k=9, weight=np.ones(normal_oracle.shape[1]))
    np.testing.assert_array_equal(ones.beta, plain.beta)
    w = np.ones(normal_oracle.shape[1]); w[:5000] = 100.0           # favour the first half of the genome
    heavy = iht.fit_iht(y, normal_oracle, None, k=9, weight=w)
    assert (np.flatnonzero(heavy.beta) < 5000).sum() > (np.flatnonzero(plain.beta) < 5000).sum()
    with pytest.raises(ValueError):
        iht.fit_iht(y, normal_oracle, None, k=9, weight=np.ones(3))


def test_project_group_sparse_properties():
    """The reference's own property tests (test/utilities_test.jl:180-213)."""
    rng = np.random.default_rng(3)
    y = rng.standard_normal(10); group = rng.integers(1, 6, 10); x = y.copy()
    iht.project_group_sparse(x, group, 2, 3)
    nz = np.flatnonzero(x)
    assert np.all(x[nz] == y[nz]) and nz.size <= 6
    y = 5 * rng.random(15); yc = y.copy()
    group = np.array([1, 2, 2, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 5])
    iht.project_group_sparse(y, group, 2, [1, 1, 2, 2, 3])
    nz = np.flatnonzero(y)
    assert np.all(y[nz] == yc[nz]) and nz.size <= 2 + 3
    y = rng.standard_normal(100000); x = y.copy()
    iht.project_k(x, 10)
    iht.project_group_sparse(y, np.ones(100000, dtype=int), 1, 10)
    assert np.all(x == y)


def test_grouped_fit_selects_J_times_k(normal_data, normal_oracle):
    """test/L0_reg_test.jl:236-242: J groups with k predictors each give J*k non-zero coefficients."""
    p = normal_oracle.shape[1]
    group = np.arange(p) // 500 + 1
    res = iht.fit_iht(normal_data["y"], normal_oracle, None, k=3, J=3, group=group)
    nz = np.flatnonzero(res.beta)
    assert nz.size == 9 and len(set(group[nz])) == 3
    with pytest.raises(ValueError):
        iht.fit_iht(normal_data["y"], normal_oracle, None, k=[600] * 20, J=2, group=group)


def test_oracle_regression_next_tier_options(normal_data, normal_oracle):
    """The oracle's frozen answers for weights / debias / groups / init_beta on the bundled data
    (tests/golden/make_oracle_regression.py; oracle-derived, not reference-published)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("mk", os.path.join(GOLDEN, "make_oracle_regression.py"))
    mk = importlib.util.module_from_spec(spec); spec.loader.exec_module(mk)
    frozen = json.load(open(os.path.join(GOLDEN, "oracle_regression.json")))
    for name, kw in mk.cases(normal_oracle.shape[1]).items():
        got = mk.run(normal_data["y"], normal_oracle, kw)
        want = frozen[name]
        assert got["iter"] == want["iter"] and got["support_0based"] == want["support_0based"], name
        np.testing.assert_allclose(got["beta"], want["beta"], rtol=1e-9, err_msg=name)
        np.testing.assert_allclose(got["c"], want["c"], rtol=1e-9, err_msg=name)
        assert abs(got["logl"] - want["logl"]) <= 1e-9 * abs(want["logl"]), name
    cv = mk.cv_case(normal_data["y"], normal_oracle)
    assert cv["iters"] == frozen["cv_q3_path_3_6_9"]["iters"]
    np.testing.assert_allclose(cv["grid"], frozen["cv_q3_path_3_6_9"]["grid"], rtol=1e-9)
    np.testing.assert_allclose(cv["mse"], frozen["cv_q3_path_3_6_9"]["mse"], rtol=1e-9)


def test_cpu_baseline_simd_kernels_agree_with_scalar():
    """oracle/csrc/cpu_ref.c: the AVX2 / AVX-512 column dots of the CPU baseline equal the scalar table loop (FP64
    accumulation, different addition order) and the numpy oracle, including missing genotypes and ragged n."""
    from oracle import cpu as ocpu
    from mendeliht_jl_b200 import synth
    lib = ocpu.load()
    rng = np.random.default_rng(4)
    for n, p, miss in ((1003, 300, 0.02), (4099, 64, 0.0), (37, 9, 0.2)):
        bed = synth.packed_columns(11, n, np.arange(p), miss)
        x = ocpu.PackedSnpLinAlgCPU(bed, n)
        v = rng.normal(size=n)
        want = snp.SnpLinAlgOracle(bed, n).xt_v(v)
        top = lib.cpu_set_simd_level(-1)
        outs = []
        for level in range(top + 1):
            assert lib.cpu_set_simd_level(level) == level
            outs.append(x.xt_v(v))
        lib.cpu_set_simd_level(-1)
        for o in outs:
            np.testing.assert_allclose(o, want, rtol=0, atol=1e-11 * np.abs(want).max())
