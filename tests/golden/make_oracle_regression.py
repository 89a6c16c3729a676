"""Freezes the ORACLE's answers for the next-tier options on the bundled `normal` data set (weights, debias, groups,
init_beta) into oracle_regression.json, so that `-m "not gpu"` catches unintended changes of the restatement.
These are NOT reference outputs (no Julia here): the reference publishes no numbers for these options, its tests
assert properties only (test/L0_reg_test.jl:236-242, test/utilities_test.jl:180-213, test/cv_iht_test.jl:26-34).
Run from the repository root:  python tests/golden/make_oracle_regression.py"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import iht, snp  # noqa: E402


def cases(p):
    w = 1.0 + (np.arange(p) % 7) / 3.0
    blocks = np.arange(p) // 500 + 1
    ks = [3] * 20
    ks[6] = 1
    return {
        "weights": dict(k=9, weight=w),
        "debias": dict(k=9, debias=True),
        "groups_J3_k3": dict(k=3, J=3, group=blocks),
        "groups_ks_vector": dict(k=ks, J=4, group=blocks),
        "init_beta": dict(k=9, init_beta=True),
    }


def run(y, o, kw):
    res = iht.fit_iht(y, o, None, **kw)
    nz = np.flatnonzero(res.beta)
    return {"iter": int(res.iter), "logl": float(res.logl), "support_0based": [int(j) for j in nz],
            "beta": [float(v) for v in res.beta[nz]], "c": [float(v) for v in res.c], "sigma_g": float(res.sigma_g)}


def cv_case(y, o):
    """3-fold cross-validation over k = 3, 6, 9 with deterministic folds (oracle/cv.py)."""
    from oracle import cv as ocv
    folds = 1 + (np.arange(o.shape[0]) % 3)
    mse, grid, iters = ocv.cv_iht(y, o, None, path=[3, 6, 9], q=3, folds=folds, return_grid=True)
    return {"mse": [float(v) for v in mse], "grid": [float(v) for v in grid], "iters": [int(v) for v in iters]}


def main():
    n = 1000
    raw = np.fromfile(os.path.join(HERE, "normal.bed"), dtype=np.uint8)[3:]
    bed = raw.reshape(-1, (n + 3) // 4)
    y = np.loadtxt(os.path.join(HERE, "normal_y.txt"))
    o = snp.SnpLinAlgOracle(bed, n)
    out = {name: run(y, o, kw) for name, kw in cases(bed.shape[0]).items()}
    out["cv_q3_path_3_6_9"] = cv_case(y, o)
    with open(os.path.join(HERE, "oracle_regression.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
