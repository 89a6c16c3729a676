import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "scale: oracle parity at BASELINE.json sizes (minutes of CPU work)")
    config.addinivalue_line("markers", "multigpu: needs at least 2 CUDA devices in one box")


def _cuda_devices() -> int:
    """Devices the product library can see; 0 when the library is missing or there is no driver."""
    try:
        import mendeliht_jl_b200 as m
        return int(m.device_count())
    except Exception:
        return 0


def pytest_collection_modifyitems(config, items):
    """A plain `pytest tests/` on a CPU-only machine skips the GPU tests instead of failing them."""
    ndev = None
    for item in items:
        if "gpu" not in item.keywords:
            continue
        if ndev is None:
            ndev = _cuda_devices()
        if ndev == 0:
            item.add_marker(pytest.mark.skip(reason="no CUDA device / libihtb200.so (GPU tests run with -m gpu on a B200)"))
        elif "multigpu" in item.keywords and ndev < 2:
            item.add_marker(pytest.mark.skip(reason="needs >= 2 CUDA devices"))


def standardize_covariates(z):
    """`standardize!` with the n-1 std on every column but the intercept (reference src/utilities.jl:494-530,
    applied by `parse_covariates`, src/wrapper.jl:228-247)."""
    z = np.array(z, dtype=np.float64, copy=True)
    n = z.shape[0]
    for j in range(1, z.shape[1]):
        m = z[:, j].sum() / n
        s = 1.0 / np.sqrt(((z[:, j] - m) ** 2).sum() / (n - 1))
        z[:, j] = (z[:, j] - m) * s
    return z


@pytest.fixture(scope="session")
def normal_data():
    """Bundled reference fixture data/normal.* (1000 x 10000, no missing) + standardised covariates."""
    from oracle import snp
    y = np.loadtxt(os.path.join(GOLDEN, "normal_y.txt"))
    n = y.shape[0]
    bed = snp.read_bed(os.path.join(GOLDEN, "normal.bed"), n)
    z = standardize_covariates(np.loadtxt(os.path.join(GOLDEN, "covariates.txt"), delimiter=","))
    return {"y": y, "n": n, "bed": bed, "z": z}


@pytest.fixture(scope="session")
def normal_oracle(normal_data):
    from oracle import snp
    return snp.SnpLinAlgOracle(normal_data["bed"], normal_data["n"])
