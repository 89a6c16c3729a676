"""Measurements for the rows SURVEY.md 8(f) marks "next", on ONE GPU at the BASELINE configs[1] shape (n=50k, p=500k,
Bernoulli/Logit, k=20): pipelined .bed ingest, prior weights, debias, group projection, init_beta.
usage: python scripts/next_rows_bench.py > profiles/<round>_next_rows.jsonl"""
import ctypes as C
import json
import os
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import mendeliht_jl_b200 as m
from mendeliht_jl_b200 import synth

N = int(os.environ.get("IHTB_BENCH_N", 50_000))
P = int(os.environ.get("IHTB_BENCH_P", 500_000))
K, SEED = 20, 2024
lib = m.load()


def out(**kw):
    print(json.dumps(kw), flush=True)


# ---- ingest: host-resident PLINK columns -> HBM (tiled layout) + statistics ------------------------------------------
p_in = min(P, 200_000)
stride = (N + 3) // 4
bed = np.empty((p_in, stride), dtype=np.uint8)
m._lib.check(lib.ihtb_synth_host(N, p_in, 0, SEED, 0.001, bed.ctypes.data_as(C.POINTER(C.c_uint8))))
for rep in range(2):
    t0 = time.perf_counter()
    g_in = m.B200SnpLinAlg.from_bed_columns(bed, N)
    dt = time.perf_counter() - t0
    g_in.close()
out(row="f2 ingest (host array)", n=N, p=p_in, bytes=int(bed.nbytes), seconds=dt, gb_per_s=bed.nbytes / dt / 1e9,
    note="ihtb_geno_create: pinned double-buffered H2D + repack + mu/sigma/missing index, second call")
with tempfile.TemporaryDirectory(dir="/dev/shm" if os.path.isdir("/dev/shm") else None) as td:
    path = os.path.join(td, "x.bed")
    with open(path, "wb") as fh:
        fh.write(bytes([0x6C, 0x1B, 0x01])); fh.write(bed.tobytes())
    t0 = time.perf_counter()
    g_in = m.B200SnpLinAlg.from_bed_file(path, N)
    dt = time.perf_counter() - t0
    t1 = time.perf_counter(); maf = g_in.maf(); cnt = g_in.counts(); dt2 = time.perf_counter() - t1
    g_in.close()
out(row="f2 ingest (mmapped .bed file)", n=N, p=p_in, bytes=int(bed.nbytes), seconds=dt, gb_per_s=bed.nbytes / dt / 1e9,
    maf_and_counts_seconds=dt2)
del bed

# ---- fits on the device-generated matrix ----------------------------------------------------------------------------
g = m.B200SnpLinAlg.synthetic(N, P, SEED)
y, z, idx, beta, _ = synth.simulate_response(SEED + 1, N, P, K, "Bernoulli", geno_seed=SEED)


def fit(label, **kw):
    m.fit_iht(y, g, z, d="Bernoulli", l="LogitLink", **kw)             # warm-up (workspace cache, clocks)
    res = m.fit_iht(y, g, z, d="Bernoulli", l="LogitLink", **kw)
    nz = np.flatnonzero(res.beta)
    out(row=label, iterations=res.iter, fit_seconds=res.time, iters_per_sec=res.iter / res.time,
        sweeps=res.n_sweeps, sweep_ms=res.sweep_seconds / max(res.n_sweeps - 1, 1) * 1e3, backtracks=res.n_backtracks,
        launches=res.n_launches, nnz=int(nz.size), true_found=int(np.intersect1d(nz, idx).size), logl=res.logl)
    return res


fit("baseline fit (k=20)", k=K)
fit("f4 prior weights (maf_weights, max 5)", k=K, weight=m.maf_weights(g, 5.0))
fit("f3 debias=true", k=K, debias=True)
blocks = np.arange(P) // 1000 + 1
fit("f4 groups: 500 blocks of 1000 SNPs, J=10, k=2", k=2, J=10, group=blocks)
fit("same shape, no groups, exact sweep (for comparison)", k=K, sweep_mode=m.SWEEP_EXACT)
yn, zn, idxn, _, _ = synth.simulate_response(SEED + 2, N, P, K, "Normal", geno_seed=SEED)
for label, kw in (("Normal fit", {}), ("f1 Normal fit with init_beta", {"init_beta": True})):
    m.fit_iht(yn, g, zn, k=K, **kw)
    res = m.fit_iht(yn, g, zn, k=K, **kw)
    # the initialisation itself (init_iht_indices! with / without initialize_beta!) is outside res.time: time it apart
    v = m.IHTVariable(g, zn, yn, K)
    v.init_iht_indices(None, kw.get("init_beta", False))
    t0 = time.perf_counter()
    for _ in range(5):
        v.init_iht_indices(None, kw.get("init_beta", False))
    t_init = (time.perf_counter() - t0) / 5
    v.close()
    out(row=label, iterations=res.iter, fit_seconds=res.time, init_seconds=t_init, total_seconds=res.time + t_init,
        sweeps=res.n_sweeps, nnz=int(np.count_nonzero(res.beta)),
        true_found=int(np.intersect1d(np.flatnonzero(res.beta), idxn).size))
