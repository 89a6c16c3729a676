"""Golden answer of BASELINE configs[4] (the north-star target run): synthetic n=500k x p=1M Normal, k=100, intercept +
10 covariates, computed ONCE by the CPU oracle with a column-streamed operator (the packed matrix is 125 GB and is
regenerated chunk by chunk on every sweep; about 7 min per sweep on 8 cores).  Writes
tests/golden/northstar_500k_x_1m.json, which bench.py --gpus 8 (`north_star` object) and tests/test_gpu_multi.py
compare the 8-GPU fit against.   usage: python scripts/make_northstar_golden.py [n p k]  (defaults 500000 1000000 100)"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np

from mendeliht_jl_b200 import synth
from oracle import glm, iht
from oracle.stream import SynthStreamSnpLinAlgCPU

n = int(sys.argv[1]) if len(sys.argv) > 1 else 500_000
p = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
k = int(sys.argv[3]) if len(sys.argv) > 3 else 100
SEED, NCOV = 2027, 10
out = sys.argv[4] if len(sys.argv) > 4 else os.path.join(ROOT, "tests", "golden", f"northstar_{n}_x_{p}.json")

t0 = time.time()
y, z, true_idx, true_beta, true_c = synth.simulate_response(SEED, n, p, k, "Normal", n_cov=NCOV, geno_seed=SEED)
print(f"response simulated in {time.time() - t0:.1f}s", flush=True)
x = SynthStreamSnpLinAlgCPU(SEED, n, p, 0.0, chunk_cols=20000, verbose=True)
t0 = time.time()
res = iht.fit_iht(y, x, z, k=k, d=glm.NORMAL, l=glm.IDENTITY)
dt = time.time() - t0
nz = np.flatnonzero(res.beta)
gold = {
    "config": f"BASELINE configs[4]: synthetic PLINK n={n} p={p} Normal/IdentityLink k={k}, intercept + {NCOV} covariates",
    "generator": {"geno_seed": SEED, "response": f"synth.simulate_response({SEED}, n, p, {k}, 'Normal', n_cov={NCOV}, geno_seed={SEED})"},
    "oracle": "oracle.iht.fit_iht over oracle.stream.SynthStreamSnpLinAlgCPU (C+OpenMP kernels, column-streamed)",
    "oracle_seconds": dt, "oracle_threads": x.threads, "oracle_matrix_passes": x.passes,
    "iter": int(res.iter), "logl": float(res.logl), "sigma_g": float(res.sigma_g),
    "support": [int(j) for j in nz], "beta": [float(res.beta[j]) for j in nz], "c": [float(v) for v in res.c],
    "trace_logl": [float(v) for v in res.trace.logl], "trace_backtracks": [int(v) for v in res.trace.backtracks],
    "trace_tol": [float(v) for v in res.trace.tol],
    "true_positives": int(np.intersect1d(nz, true_idx).size),
}
with open(out, "w") as f:
    json.dump(gold, f, indent=1)
print(f"oracle fit: {res.iter} iterations, {x.passes} matrix passes in {dt:.0f}s -> {out}", flush=True)
