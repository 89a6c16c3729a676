"""BASELINE configs[4] (north-star target) driven by ONE process over every GPU of the box -- the path a Julia caller
takes (ihtb_mgeno SHARD + ihtb_mfit: one host thread per device, peer-memory collectives, no torchrun, no NCCL):
synthetic n=500k x p=1M Normal, k=100, intercept + 10 covariates; compared with the CPU oracle's golden answer.
usage: python scripts/northstar_inprocess.py [ngpu]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np

import mendeliht_jl_b200 as m
from mendeliht_jl_b200 import synth

ngpu = int(sys.argv[1]) if len(sys.argv) > 1 else m.device_count()
n, p, k, ncov, seed = 500_000, 1_000_000, 100, 10, 2027
t0 = time.perf_counter()
g = m.B200MultiSnpLinAlg.synthetic(n, p, seed, 0.0, ngpu=ngpu, mode=m.B200MultiSnpLinAlg.SHARD)
t_gen = time.perf_counter() - t0
y, z, true_idx, _, _ = synth.simulate_response(seed, n, p, k, "Normal", n_cov=ncov, geno_seed=seed)
m.fit_iht(y, g, z, k=k)                                            # warm-up: workspaces, peer mappings
times = []
for _ in range(3):
    t0 = time.perf_counter()
    res = m.fit_iht(y, g, z, k=k)
    times.append(time.perf_counter() - t0)
gold = json.load(open(os.path.join(ROOT, "tests", "golden", f"northstar_{n}_x_{p}.json")))
nz = np.flatnonzero(res.beta)
gs = np.asarray(gold["support"])
same = bool(np.array_equal(nz, gs))
out = {"config": gold["config"] + f", one process driving {ngpu} GPUs (ihtb_mfit)", "n_gpus": ngpu,
       "iterations": int(res.iter), "oracle_iterations": gold["iter"], "support_identical": same,
       "backtracks_identical": [t[1] for t in res.trace] == gold["trace_backtracks"],
       "max_rel_err_beta": float(np.max(np.abs(res.beta[gs] - np.asarray(gold["beta"])) / np.abs(gold["beta"]))) if same else None,
       "rel_err_logl": abs(res.logl - gold["logl"]) / abs(gold["logl"]),
       "max_rel_err_logl_trace": float(np.max(np.abs(np.array([t[0] for t in res.trace]) - np.asarray(gold["trace_logl"])) / np.abs(gold["trace_logl"]))),
       "call_seconds_best_of_3": min(times), "fit_seconds_in_library": res.time, "iterations_per_sec": res.iter / res.time,
       "sweep_ms": res.sweep_seconds / max(res.n_sweeps - 1, 1) * 1e3, "generate_s": t_gen,
       "true_positives": int(np.intersect1d(nz, true_idx).size)}
out["oracle_parity"] = bool(same and out["iterations"] == out["oracle_iterations"] and out["backtracks_identical"]
                            and out["max_rel_err_beta"] <= 1e-6 and out["rel_err_logl"] <= 1e-6)
print(json.dumps(out))
g.close()
