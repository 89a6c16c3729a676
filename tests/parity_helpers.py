"""Shared assertions of the GPU parity tests: the CUDA fit (through the C ABI) against the oracle's fit on the same
inputs.  Bar (BASELINE.json north_star): support / iteration / backtrack counts identical, values within 1e-6."""
import numpy as np

RTOL = 1e-6


def compare_fit(res, ref, rtol=RTOL, check_backtracks=True):
    assert res.iter == ref.iter
    assert np.array_equal(np.flatnonzero(res.beta), np.flatnonzero(ref.beta))      # support: exact
    np.testing.assert_allclose(res.beta, ref.beta, rtol=rtol, atol=1e-12)
    np.testing.assert_allclose(res.c, ref.c, rtol=rtol, atol=1e-12)
    if np.isfinite(ref.logl):
        assert abs(res.logl - ref.logl) <= rtol * abs(ref.logl)
    else:
        assert res.logl == ref.logl
    assert abs(res.sigma_g - ref.sigma_g) <= rtol * abs(ref.sigma_g) + 1e-12
    if check_backtracks:
        assert [t[1] for t in res.trace] == ref.trace.backtracks
        np.testing.assert_allclose([t[2] for t in res.trace], ref.trace.tol, rtol=max(1e-5, 10 * rtol), atol=1e-12)
    np.testing.assert_allclose([t[0] for t in res.trace], ref.trace.logl, rtol=rtol)


def compare_mv_fit(res, ref, rtol=RTOL):
    assert res.iter == ref.iter
    assert np.array_equal(res.beta != 0, ref.beta != 0)
    np.testing.assert_allclose(res.beta, ref.beta, rtol=rtol, atol=1e-12)
    np.testing.assert_allclose(res.c, ref.c, rtol=rtol, atol=1e-12)
    assert abs(res.logl - ref.logl) <= rtol * abs(ref.logl)
    np.testing.assert_allclose(res.Sigma, ref.Sigma, rtol=1e-5, atol=1e-10)
    np.testing.assert_allclose(res.sigma_g, ref.sigma_g, rtol=rtol)
    assert [t[1] for t in res.trace] == ref.trace.backtracks
    np.testing.assert_allclose([t[0] for t in res.trace], ref.trace.logl, rtol=rtol)


