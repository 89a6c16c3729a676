"""Where the public-API call spends its time on a sharded fit: torchrun --nproc-per-node N scripts/e2e_probe.py
(configs[1] shape per GPU).  Prints rank 0's wall-clock per stage of api.fit_iht, averaged over the fits."""
import json
import os
import sys
import time

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
n, k = 50_000, 20
p = 500_000 * world
j0, pl = parallel.shard_range(p, world, rank)
g = m.B200SnpLinAlg.synthetic(n, pl, 2024, 0.0, j0)
y, z, *_ = synth.simulate_response(2025, n, p, k, "Bernoulli", geno_seed=2024)
T = {"create": 0.0, "init": 0.0, "fit": 0.0, "get": 0.0, "close": 0.0, "whole_call": 0.0}
for rep in range(6):
    t0 = time.perf_counter()
    v = m.IHTVariable(g, z, y, k, "Bernoulli", "LogitLink", comm=comm, p_global=p)
    t1 = time.perf_counter()
    v.init_iht_indices(None)
    t2 = time.perf_counter()
    res, tr = v.fit()
    t3 = time.perf_counter()
    beta, c, _, _ = v.get()
    t4 = time.perf_counter()
    v.close()
    t5 = time.perf_counter()
    r = m.fit_iht(y, g, z, k=k, d="Bernoulli", l="LogitLink", comm=comm, p_global=p)
    t6 = time.perf_counter()
    if rep:
        for key, dt in zip(T, (t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4, t6 - t5)):
            T[key] += dt / 5
if rank == 0:
    print(json.dumps({"world": world, "iterations": int(res.iter), "library_run_ms": res.time * 1e3,
                      "ms": {k_: round(v_ * 1e3, 3) for k_, v_ in T.items()}}))
dist.barrier(); comm.close(); dist.destroy_process_group()
