"""torchrun --nproc-per-node N scripts/allreduce_latency.py : device time per all-reduce of an n-vector of doubles,
ncclAllReduce vs the peer-memory push + local reduce (p2p.cu).  Run on a multi-GPU box."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from mendeliht_jl_b200 import parallel

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
comm = parallel.Comm(dist, lr)
rows = []
for n in (500000, 50000, 5000):     # largest first: the first call fixes the peer slot capacity
    for p2p in (False, True):
        us = comm.allreduce_latency_us(n, 300, p2p)
        t = torch.tensor([us], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        rows.append({"n_doubles": n, "path": "p2p" if p2p else "nccl", "us_max_over_ranks": round(t.item(), 2)})
if rank == 0:
    print(json.dumps({"n_gpus": world, "allreduce": rows}))
dist.barrier()
comm.close()
dist.destroy_process_group()
