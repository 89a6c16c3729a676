"""Host wall-clock per phase of a (possibly SNP-sharded) fit at north-star shape: n = 500k samples, 125k columns per GPU,
Normal, k = 100, intercept + 10 covariates.  usage: [torchrun --nproc-per-node N] python scripts/phase_probe.py [n p_per_gpu k ncov]"""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import mendeliht_jl_b200 as m
from mendeliht_jl_b200 import parallel, synth

a = [int(x) for x in sys.argv[1:]]
n, ppg, k, ncov = (a + [500_000, 125_000, 100, 10][len(a):])[:4]
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); lr = int(os.environ.get("LOCAL_RANK", "0"))
comm = None
if world > 1:
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    comm = parallel.Comm(dist, lr)
m._lib.check(m.load().ihtb_set_device(lr))
p = ppg * world
j0, pl = parallel.shard_range(p, world, rank)
g = m.B200SnpLinAlg.synthetic(n, pl, 2027, 0.0, j0)
y, z, *_ = synth.simulate_response(2027, n, p, k, "Normal", n_cov=ncov, geno_seed=2027)
v = m.IHTVariable(g, z, y, k, "Normal", "IdentityLink", comm=comm, p_global=p)
for rep in range(3):
    ph0 = (C.c_double * 4)(); m._lib.check(m.load().ihtb_fit_phase_times(v._h, ph0))
    v.init_iht_indices(None)
    res, tr = v.fit()
    ph1 = (C.c_double * 4)(); m._lib.check(m.load().ihtb_fit_phase_times(v._h, ph1))
    if rank == 0 and rep == 2:
        it = max(int(res.n_steps), 1)
        print(json.dumps({"n": n, "p": p, "world": world, "iterations": int(res.iter), "steps": int(res.n_steps),
                          "run_ms": res.time * 1e3, "sweep_ms_each": res.sweep_seconds / max(res.n_sweeps - 1, 1) * 1e3,
                          "per_iteration_ms": {nm: (ph1[i] - ph0[i]) / it * 1e3 for i, nm in
                                               enumerate(["stepsize", "gradstep", "xb_glm", "score_sweep"])},
                          "backtracks": int(res.n_backtracks), "launches_per_iteration": res.n_launches / it}))
v.close()
if comm is not None:
    dist.barrier(); comm.close(); dist.destroy_process_group()
