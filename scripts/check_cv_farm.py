"""torchrun --nproc-per-node N scripts/check_cv_farm.py : the (fold, k) cross-validation grid farmed over N GPUs
(replicas of the matrix, no data-path collective) equals the single-GPU cv_iht; prints the speed-up."""
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
m._lib.check(m.load().ihtb_set_device(lr))
n, p, q = int(os.environ.get("CV_N", 20000)), int(os.environ.get("CV_P", 100000)), 5
path = list(range(1, 21))
g = m.B200SnpLinAlg.synthetic(n, p, 2025)
y, z, *_ = synth.simulate_response(2025, n, p, 10, "Poisson", geno_seed=2025)
folds = synth.folds_for(2025, n, q)
grid = m.allocate_fold_and_k(q, path)


def cv_fn(combos):
    return m.cv_iht(y, g, z, d="Poisson", l="LogLink", path=path, q=q, folds=folds, combos=combos, return_grid=True)


dist.barrier(); torch.cuda.synchronize()
t0 = time.perf_counter()
mses, iters = parallel.cv_iht_farm(dist, cv_fn, len(grid))
dist.barrier()
t_farm = time.perf_counter() - t0
if rank == 0:
    t0 = time.perf_counter()
    ref_m, ref_i = m.cv_iht(y, g, z, d="Poisson", l="LogLink", path=path, q=q, folds=folds, return_grid=True)
    t_one = time.perf_counter() - t0
    ok = np.array_equal(iters, ref_i) and np.allclose(mses, ref_m, rtol=1e-12, atol=0)
    mse = m.meanloss(mses, q, folds)
    print(f"cv farm on {world} GPUs: {t_farm:.2f}s vs {t_one:.2f}s on one GPU (x{t_one / t_farm:.2f}); identical={ok}; "
          f"best k={path[int(np.argmin(mse))]}; total iterations={int(iters.sum())}", flush=True)
dist.barrier()
dist.destroy_process_group()
