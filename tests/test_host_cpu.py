"""CPU-only checks of the host layer: the C-ABI library loads and exports every symbol include/ihtb200.h declares,
the numpy twin of the device generator is deterministic, and the product never imports the oracle."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "ihtb200.h")).read()
    return sorted(set(re.findall(r"^\s*int32_t\s+(ihtb_\w+)\s*\(", hdr, flags=re.M)))


def test_library_exports_every_declared_symbol():
    import mendeliht_jl_b200 as m
    lib = m.load()
    syms = _declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/ihtb200.h but not exported"
    # and the ctypes table covers exactly the header
    assert sorted(m._lib.SIGNATURES) == syms
    assert lib.ihtb_version() >= 100


def test_struct_layouts_match_header():
    import mendeliht_jl_b200 as m
    assert ctypes.sizeof(m._lib.Cfg) == 56
    assert ctypes.sizeof(m._lib.Result) == 72
    assert ctypes.sizeof(m._lib.IterTrace) == 32


def test_no_cpu_fallback_without_device():
    import mendeliht_jl_b200 as m
    if m.device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(m.CudaError):
        m.B200SnpLinAlg.synthetic(64, 8, 1)
    with pytest.raises(m.CudaError):
        m.B200SnpLinAlg.from_bed_columns(np.zeros((8, 16), dtype=np.uint8), 64)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "mendeliht.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, f


def test_synth_numpy_twin_is_deterministic_and_sane():
    from mendeliht_jl_b200 import synth
    a = synth.codes(2024, 2000, np.arange(50))
    b = synth.codes(2024, 2000, np.arange(50))
    assert np.array_equal(a, b)
    assert set(np.unique(a)) <= {0, 2, 3}
    # columns are a pure function of (seed, j): any subset reproduces
    assert np.array_equal(synth.codes(2024, 2000, [7, 31]), a[:, [7, 31]])
    # allele frequency follows maf_j
    _, maf = synth.maf_threshold(synth.col_key(2024, np.arange(50)))
    dos = np.array([0, 0, 1, 2])[a]
    assert np.all(np.abs(dos.mean(axis=0) / 2 - maf) < 0.05)
    m = synth.codes(2024, 4000, np.arange(20), missing_rate=0.05)
    assert 0.03 < (m == 1).mean() < 0.07
    pk = synth.packed_columns(2024, 1003, [0, 1, 2])
    from oracle import snp
    assert np.array_equal(snp.unpack_codes(pk, 1003), synth.codes(2024, 1003, [0, 1, 2]))


def test_cv_grid_and_meanloss_match_oracle():
    import mendeliht_jl_b200 as m
    from oracle import cv as ocv
    assert m.allocate_fold_and_k(5, range(1, 21)) == ocv.allocate_fold_and_k(5, range(1, 21))
    rng = np.random.default_rng(0)
    folds = rng.integers(1, 6, size=1000)
    loss = rng.random(100)
    np.testing.assert_array_equal(m.meanloss(loss, 5, folds), ocv.meanloss(loss, 5, folds))


def test_argument_checks_mirror_reference_asserts():
    import mendeliht_jl_b200 as m
    from mendeliht_jl_b200 import api
    with pytest.raises(AssertionError):
        api._check_args(k=-1, max_iter=10, max_step=3, tol=1e-4)
    with pytest.raises(AssertionError):
        api._check_args(k=1, max_iter=-1, max_step=3, tol=1e-4)
    with pytest.raises(AssertionError):
        api._check_args(k=1, max_iter=10, max_step=3, tol=1e-17)
    assert m.canonicallink("NegativeBinomial") == "LogLink" and m.canonicallink("Bernoulli") == "LogitLink"


def test_wrapper_parsers_without_gpu(tmp_path):
    """`parse_covariates` (src/wrapper.jl:228-247) and the .fam phenotype reader (:170-191) are pure host code."""
    import mendeliht_jl_b200 as m
    from mendeliht_jl_b200 import api
    from conftest import GOLDEN
    z = m.parse_covariates(os.path.join(GOLDEN, "covariates.txt"))
    raw = np.loadtxt(os.path.join(GOLDEN, "covariates.txt"), delimiter=",")
    assert np.all(z[:, 0] == 1)                                   # the intercept column is never standardised
    np.testing.assert_allclose(z[:, 1].mean(), 0.0, atol=1e-12)
    np.testing.assert_allclose(z[:, 1].std(ddof=1), 1.0, rtol=1e-12)
    np.testing.assert_allclose(z[:, 1], (raw[:, 1] - raw[:, 1].mean()) / raw[:, 1].std(ddof=1), rtol=1e-12)
    z_excl = m.parse_covariates(os.path.join(GOLDEN, "covariates.txt"), exclude_std_idx=[2])
    np.testing.assert_array_equal(z_excl[:, 1], raw[:, 1])
    fam = tmp_path / "t.fam"
    fam.write_text("f1 i1 0 0 1 1.5\nf2 i2 0 0 2 -9\nf3 i3 0 0 1 NA\nf4 i4 0 0 2 2.5\n")
    y = api._read_fam_phenotype(str(fam))
    np.testing.assert_allclose(y, [1.5, 2.0, 2.0, 2.5])          # missing -> mean of the observed
    # binary / count traits cannot be imputed: the reference throws (src/wrapper.jl:193-207)
    for d in ("Bernoulli", "Poisson", "NegativeBinomial"):
        with pytest.raises(api.MissingPhenotype):
            api._read_fam_phenotype(str(fam), 6, d)
    fam2 = tmp_path / "t2.fam"
    fam2.write_text("f1 i1 0 0 1 1.5 4\nf2 i2 0 0 2 -9 6\nf3 i3 0 0 1 0.5 NA\n")
    np.testing.assert_allclose(api._read_fam_phenotypes_mv(str(fam2), [6, 7]), [[1.5, 1.0, 0.5], [4.0, 6.0, 5.0]])
    w = np.array([0.5, 0.1, 0.01])
    assert m.canonicallink("NegativeBinomial") == "LogLink" and m.canonicallink("Bernoulli") == "LogitLink"
    assert m.allocate_fold_and_k(2, [3, 5]) == [(1, 3), (1, 5), (2, 3), (2, 5)]
    with pytest.raises(AssertionError):
        api._check_args(5, 10, 3, 0.0)


def test_ternary_tile_format_round_trip():
    """synth.ternary_tiles restates the byte stream the sweeps read (five base-3 dosages per byte, 640-sample slabs,
    columns of a quad permuted by word position): every sample decodes back to its dosage, missing to 0."""
    import numpy as np
    from mendeliht_jl_b200 import synth
    for n, p, miss in [(640, 8, 0.0), (1003, 13, 0.05), (1281, 5, 0.3)]:
        bed = synth.packed_columns(9, n, np.arange(p), miss)
        t = synth.ternary_tiles(bed, n)
        slabs, p4 = -(-n // 640), (p + 3) // 4 * 4
        assert t.dtype == np.uint8 and t.shape == (slabs * p4 * 128,) and t.max() <= 242
        T = t.reshape(slabs, p4 // 4, 32, 4, 4)
        dos = np.array([0, 0, 1, 2])[synth.codes(9, n, np.arange(p), miss)]              # [n, p]
        i = np.arange(n)
        s, r = np.divmod(i, 640); w, rr = np.divmod(r, 20); b, dg = np.divmod(rr, 5)
        for j in range(p):
            byte = T[s, j // 4, w, (j & 3) ^ (w & 3), b].astype(np.int64)
            assert np.array_equal((byte // 3 ** dg) % 3, dos[:, j])
        assert not T[:, :, :, :, :].reshape(slabs, p4 // 4, 32, 4, 4)[:, p // 4:, :, :, :].any() or p % 4      # padding columns are zero


def test_ihtresult_builds_dense_beta_on_first_access():
    """The fit hands the model over as k (index, value) pairs; `IHTResult.beta` (the reference's dense vector,
    src/data_structures.jl:245-258) is materialised on first access and cached; a dense array passed in is kept as is."""
    import pickle
    import numpy as np
    from mendeliht_jl_b200.api import IHTResult, SparseCoef
    sp = SparseCoef(10, np.array([2, 5]), np.array([1.5, -2.0]))
    r = IHTResult(0.1, -1.0, 3, sp, np.ones(1), 1, 2, [], "Normal", 0.5)
    assert "beta" not in r.__dict__ and r.beta_sparse is sp
    want = np.zeros(10); want[[2, 5]] = [1.5, -2.0]
    assert np.array_equal(r.beta, want) and r.beta is r.beta and "beta" in r.__dict__
    assert np.array_equal(pickle.loads(pickle.dumps(r)).beta, want)
    dense = IHTResult(0.1, -1.0, 3, want.copy(), np.ones(1), 1, 2, [], "Normal", 0.5)
    assert dense.beta_sparse is None and np.array_equal(dense.beta, want)
    try:
        dense.no_such_field
        raise AssertionError("AttributeError expected")
    except AttributeError:
        pass
