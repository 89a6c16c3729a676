"""Runs the other BASELINE.json configs on one GPU (informational; bench.py's headline line is configs[1]).
usage: python scripts/run_configs.py [cv|mv|mvshard|ukb] ...   (mvshard: configs[3] with the SNP columns split over all visible GPUs)"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import mendeliht_jl_b200 as m
from mendeliht_jl_b200 import synth

which = sys.argv[1:] or ["mv", "cv", "ukb"]


def sweep_bytes(n, p):
    return p * ((n + 3) // 4) + 8 * n + 24 * p


if "mv" in which:      # configs[3]: MvNormal r=5 traits n=100k p=500k k=50
    n, p, r, k = 100_000, 500_000, 5, 50
    g = m.B200SnpLinAlg.synthetic(n, p, 2026)
    rng = np.random.default_rng(2026)
    idx = np.sort(rng.permutation(p)[:k])
    B = np.zeros((r, k))
    for c in range(k):
        B[rng.integers(0, r), c] = rng.normal()
    A = rng.normal(size=(r, r)); cov = A @ A.T / r + 0.5 * np.eye(r)
    Y = np.linalg.cholesky(cov) @ rng.normal(size=(r, n)) + 1.0
    for s in range(0, k, 10):
        Y += B[:, s:s + 10] @ synth.standardized_columns(2026, n, idx[s:s + 10]).T
    mode = {"fast": m.SWEEP_FAST, "pair": m.SWEEP_PAIR, "exact": m.SWEEP_EXACT}[os.environ.get("MV_SWEEP", "pair")]
    m.fit_iht(Y, g, None, k=k, sweep_mode=mode)                                   # warm-up
    t0 = time.perf_counter()
    res = m.fit_iht(Y, g, None, k=k, sweep_mode=mode)
    dt = time.perf_counter() - t0
    nz = np.flatnonzero((res.beta != 0).any(axis=0))
    print(json.dumps({"config": "configs[3] MvNormal r=5 n=100k p=500k k=50", "sweep_mode": os.environ.get("MV_SWEEP", "pair"),
                      "iterations": res.iter, "seconds": dt,
                      "fit_seconds": res.time, "iters_per_sec": res.iter / res.time, "sweeps": res.n_sweeps,
                      "sweep_seconds": res.sweep_seconds,
                      "sweep_gbs_per_rhs": res.n_sweeps * sweep_bytes(n, p) / res.sweep_seconds / 1e9 if res.sweep_seconds else None,
                      "entries": int(np.count_nonzero(res.beta)), "true_cols_found": int(np.intersect1d(nz, idx).size),
                      "logl": res.logl}))
    g.close()

if "mvshard" in which:      # configs[3] with the SNP columns split over every visible GPU, against the one-GPU fit
    n, p, r, k = 100_000, 500_000, 5, 50
    ngpu = m.device_count()
    rng = np.random.default_rng(2026)
    idx = np.sort(rng.permutation(p)[:k])
    B = np.zeros((r, k))
    for c in range(k):
        B[rng.integers(0, r), c] = rng.normal()
    A = rng.normal(size=(r, r)); cov = A @ A.T / r + 0.5 * np.eye(r)
    Y = np.linalg.cholesky(cov) @ rng.normal(size=(r, n)) + 1.0
    for s in range(0, k, 10):
        Y += B[:, s:s + 10] @ synth.standardized_columns(2026, n, idx[s:s + 10]).T
    out = {}
    for name, make in (("one_gpu", lambda: m.B200SnpLinAlg.synthetic(n, p, 2026)),
                       ("sharded", lambda: m.B200MultiSnpLinAlg.synthetic(n, p, 2026, 0.0, ngpu=ngpu,
                                                                         mode=m.B200MultiSnpLinAlg.SHARD))):
        g = make()
        m.fit_iht(Y, g, None, k=k, sweep_mode=m.SWEEP_PAIR)                       # warm-up
        best = None
        for _ in range(3):
            res = m.fit_iht(Y, g, None, k=k, sweep_mode=m.SWEEP_PAIR)
            best = res if best is None or res.time < best.time else best
        out[name] = best
        g.close()
    a, b_ = out["one_gpu"], out["sharded"]
    print(json.dumps({"config": "configs[3] MvNormal r=5 n=100k p=500k k=50, PAIR sweep", "n_gpus": ngpu,
                      "one_gpu_fit_seconds": a.time, "sharded_fit_seconds": b_.time, "speedup": a.time / b_.time,
                      "iterations": [a.iter, b_.iter], "support_identical": bool(np.array_equal(a.beta != 0, b_.beta != 0)),
                      "max_abs_diff_beta": float(np.max(np.abs(a.beta - b_.beta))),
                      "rel_diff_logl": float(abs(a.logl - b_.logl) / abs(a.logl)),
                      "sweep_seconds": [a.sweep_seconds, b_.sweep_seconds]}))

if "cv" in which:      # configs[2]: Poisson CV q=5 path 1:20 n=100k p=500k (100 fits, one GPU here)
    n, p, q = 100_000, 500_000, 5
    g = m.B200SnpLinAlg.synthetic(n, p, 2025)
    y, z, idx, beta, _ = synth.simulate_response(2025, n, p, 10, "Poisson", geno_seed=2025)
    folds = synth.folds_for(2025, n, q)
    t0 = time.perf_counter()
    # the whole grid in ONE library call (ihtb_cv_run): two fits at a time share their sweeps (IHTB_CV_PAIR=0: one at a time)
    m.cv_run(y, g, z, folds, q, [1, 2], d="Poisson", l="LogLink")                  # warm-up (workspaces)
    t0 = time.perf_counter()
    mses, iters = m.cv_run(y, g, z, folds, q, list(range(1, 21)), d="Poisson", l="LogLink")
    dt = time.perf_counter() - t0
    mse = m.meanloss(mses, q, folds)
    print(json.dumps({"config": "configs[2] Poisson CV q=5 path=1:20 n=100k p=500k", "fits": 100, "seconds": dt,
                      "pair_sweeps": os.environ.get("IHTB_CV_PAIR", "1") != "0",
                      "total_iterations": int(iters.sum()), "iters_per_sec": float(iters.sum()) / dt,
                      "best_k": int(np.argmin(mse)) + 1, "mse": [float(v) for v in mse],
                      "grid_checksum": float(np.sum(mses * np.arange(1, mses.size + 1)))}))
    g.close()

if "ukb" in which:     # configs[4]: n=500k p=1M Normal k=100 with 10 covariates, whole matrix on ONE GPU (125 GB)
    n, p, k = 500_000, 1_000_000, 100
    t0 = time.perf_counter()
    g = m.B200SnpLinAlg.synthetic(n, p, 2027)
    tg = time.perf_counter() - t0
    y, z, idx, beta, _ = synth.simulate_response(2027, n, p, k, "Normal", n_cov=10, geno_seed=2027)
    t0 = time.perf_counter()
    res = m.fit_iht(y, g, z, k=k)
    dt = time.perf_counter() - t0
    nz = np.flatnonzero(res.beta)
    print(json.dumps({"config": "configs[4] n=500k p=1M Normal k=100 q=11 on ONE B200 (125 GB packed)",
                      "generate_seconds": tg, "iterations": res.iter, "seconds": dt, "fit_seconds": res.time,
                      "iters_per_sec": res.iter / res.time, "sweeps": res.n_sweeps, "sweep_seconds": res.sweep_seconds,
                      "sweep_gbs": res.n_sweeps and (res.n_sweeps - 1) * sweep_bytes(n, p) / res.sweep_seconds / 1e9,
                      "support": int(nz.size), "true_positives": int(np.intersect1d(nz, idx).size), "logl": res.logl}))
    g.close()
