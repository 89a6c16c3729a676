"""torchrun --nproc-per-node N scripts/allreduce_latency.py : device time per all-reduce of an n-vector of doubles,
ncclAllReduce vs the two peer-memory forms of p2p.cu (push-all: every rank stores its vector into every rank, one
local reduce; two-phase: reduce-scatter + all-gather kernels).  Run on a multi-GPU box."""
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
names = {0: "nccl", 2: "p2p push-all", 3: "p2p two-phase"}
for n in (2000000, 500000, 200000, 50000, 5000):     # largest first: the first call fixes the mapped area
    for mode in (0, 2, 3):
        if mode == 2 and n > 262144:
            continue                                  # longer than a push-all slot
        us = comm.allreduce_latency_us(n, 300, mode)
        t = torch.tensor([us], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        rows.append({"n_doubles": n, "path": names[mode], "us_max_over_ranks": round(t.item(), 2)})
if rank == 0:
    print(json.dumps({"n_gpus": world, "allreduce": rows}))
dist.barrier()
comm.close()
dist.destroy_process_group()
