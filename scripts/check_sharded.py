"""torchrun --nproc-per-node N scripts/check_sharded.py : SNP-sharded fit == single-GPU fit (same support, iterations,
beta to 1e-9), for several distributions.  Run on a multi-GPU box."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

import mendeliht_jl_b200 as m
from mendeliht_jl_b200 import parallel, synth

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
comm = parallel.Comm(dist, lr)
ok = True
for d, l, n, p, k, ncov, miss in [("Normal", "IdentityLink", 5000, 20001, 8, 2, 0.0),
                                  ("Bernoulli", "LogitLink", 6000, 16000, 6, 0, 0.001),
                                  ("Poisson", "LogLink", 4000, 12000, 6, 1, 0.0)]:
    seed = 11 + n
    y, z, *_ = synth.simulate_response(seed, n, p, k, d, n_cov=ncov, missing_rate=miss)
    j0, pl = parallel.shard_range(p, world, rank)
    g_loc = m.B200SnpLinAlg.synthetic(n, pl, seed, miss, j0)
    res = m.fit_iht(y, g_loc, z, k=k + 2, d=d, l=l, comm=comm, p_global=p)
    g_full = m.B200SnpLinAlg.synthetic(n, p, seed, miss, 0)
    ref = m.fit_iht(y, g_full, z, k=k + 2, d=d, l=l)
    same = (res.iter == ref.iter and np.array_equal(np.flatnonzero(res.beta), np.flatnonzero(ref.beta))
            and np.allclose(res.beta, ref.beta, rtol=1e-9, atol=1e-12) and abs(res.logl - ref.logl) < 1e-9 * abs(ref.logl)
            and [t[1] for t in res.trace] == [t[1] for t in ref.trace])
    ok = ok and same
    print(f"rank {rank} {d}: sharded iter={res.iter} full iter={ref.iter} same={same} "
          f"max|dbeta|={np.abs(res.beta - ref.beta).max():.3e} logl {res.logl:.9f} vs {ref.logl:.9f}", flush=True)
# prior weights on a sharded fit: every rank holds the whole weight vector, the device keys use the local slice
n, p, k = 4000, 16000, 6
y, z, *_ = synth.simulate_response(55, n, p, k, "Normal", n_cov=1)
j0, pl = parallel.shard_range(p, world, rank)
g_loc = m.B200SnpLinAlg.synthetic(n, pl, 55, 0.0, j0)
g_full = m.B200SnpLinAlg.synthetic(n, p, 55, 0.0, 0)
w = m.maf_weights(g_full, max_weight=4.0)
res = m.fit_iht(y, g_loc, z, k=k + 2, weight=w, comm=comm, p_global=p)
ref = m.fit_iht(y, g_full, z, k=k + 2, weight=w)
same = (res.iter == ref.iter and np.array_equal(np.flatnonzero(res.beta), np.flatnonzero(ref.beta))
        and np.allclose(res.beta, ref.beta, rtol=1e-9, atol=1e-12))
ok = ok and same
print(f"rank {rank} weighted: sharded iter={res.iter} full iter={ref.iter} same={same}", flush=True)
res = m.fit_iht(y, g_loc, z, k=k + 2, init_beta=True, comm=comm, p_global=p)
ref = m.fit_iht(y, g_full, z, k=k + 2, init_beta=True)
same = (res.iter == ref.iter and np.array_equal(np.flatnonzero(res.beta), np.flatnonzero(ref.beta))
        and np.allclose(res.beta, ref.beta, rtol=1e-9, atol=1e-12) and np.allclose(res.c, ref.c, rtol=1e-9))
ok = ok and same
print(f"rank {rank} init_beta: sharded iter={res.iter} full iter={ref.iter} same={same}", flush=True)
blocks = np.arange(p) // 333 + 1            # groups that straddle the shard boundary
for kw in ({"k": 2, "J": 4}, {"k": [2] * int(blocks.max()), "J": 3}):
    res = m.fit_iht(y, g_loc, z, group=blocks, comm=comm, p_global=p, **kw)
    ref = m.fit_iht(y, g_full, z, group=blocks, **kw)
    same = (res.iter == ref.iter and np.array_equal(np.flatnonzero(res.beta), np.flatnonzero(ref.beta))
            and np.allclose(res.beta, ref.beta, rtol=1e-9, atol=1e-12))
    ok = ok and same
    print(f"rank {rank} groups {'ks' if not np.isscalar(kw['k']) else 'k'}: sharded iter={res.iter} full iter={ref.iter} "
          f"same={same} nnz={np.count_nonzero(res.beta)}", flush=True)
yb, zb, *_ = synth.simulate_response(56, n, p, k, "Bernoulli", geno_seed=55)
res = m.fit_iht(yb, g_loc, zb, k=k + 1, d="Bernoulli", l="LogitLink", debias=True, comm=comm, p_global=p)
ref = m.fit_iht(yb, g_full, zb, k=k + 1, d="Bernoulli", l="LogitLink", debias=True)
same = (res.iter == ref.iter and np.array_equal(np.flatnonzero(res.beta), np.flatnonzero(ref.beta))
        and np.allclose(res.beta, ref.beta, rtol=1e-8, atol=1e-12))
ok = ok and same
print(f"rank {rank} debias: sharded iter={res.iter} full iter={ref.iter} same={same} "
      f"max|dbeta|={np.abs(res.beta - ref.beta).max():.3e}", flush=True)
# multivariate (MvNormal) fit with the columns split over the ranks: PAIR and FAST sweeps, init_beta
r, kk = 3, 9
rng = np.random.default_rng(57)
idx = np.sort(rng.permutation(p)[:kk])
Bt = np.zeros((r, kk))
for cc in range(kk):
    Bt[rng.integers(0, r), cc] = rng.normal() * 0.8
Y = Bt @ synth.standardized_columns(55, n, idx).T + rng.normal(size=(r, n))
Z = np.vstack([np.ones(n), rng.normal(size=n)])
for kw in ({"sweep_mode": m.SWEEP_PAIR}, {"sweep_mode": m.SWEEP_FAST}, {"sweep_mode": m.SWEEP_PAIR, "init_beta": True}):
    res = m.fit_iht(Y, g_loc, Z, k=kk + 2, comm=comm, p_global=p, **kw)
    ref = m.fit_iht(Y, g_full, Z, k=kk + 2, **kw)
    same = (res.iter == ref.iter and np.array_equal(res.beta != 0, ref.beta != 0)
            and np.allclose(res.beta, ref.beta, rtol=1e-9, atol=1e-12) and np.allclose(res.c, ref.c, rtol=1e-9)
            and abs(res.logl - ref.logl) < 1e-9 * abs(ref.logl))
    ok = ok and same
    print(f"rank {rank} MvNormal {kw}: sharded iter={res.iter} full iter={ref.iter} same={same}", flush=True)
if os.environ.get("CHECK_SHARDED_SKIP_FULL") == "1":
    dist.barrier()
    comm.close()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)
# full BASELINE size per GPU: FAST and EXACT sweeps must give the same support / iterations / beta (stress for the pipeline)
n, pg, k = 50000, 500000 * world, 20
y, z, *_ = synth.simulate_response(2025, n, pg, k, "Bernoulli", geno_seed=2024)
j0, pl = parallel.shard_range(pg, world, rank)
g_loc = m.B200SnpLinAlg.synthetic(n, pl, 2024, 0.0, j0)
ra = m.fit_iht(y, g_loc, z, k=k, d="Bernoulli", l="LogitLink", comm=comm, p_global=pg, sweep_mode=m.SWEEP_FAST)
rb = m.fit_iht(y, g_loc, z, k=k, d="Bernoulli", l="LogitLink", comm=comm, p_global=pg, sweep_mode=m.SWEEP_EXACT)
same = (ra.iter == rb.iter and np.array_equal(np.flatnonzero(ra.beta), np.flatnonzero(rb.beta))
        and np.allclose(ra.beta, rb.beta, rtol=1e-9, atol=1e-12))
ok = ok and same
print(f"rank {rank} full-size FAST vs EXACT: iter {ra.iter}/{rb.iter} same={same}", flush=True)
dist.barrier()
comm.close()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
