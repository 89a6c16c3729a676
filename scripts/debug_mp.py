import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
stage = sys.argv[1]
import mendeliht_jl_b200 as m
if stage >= "b":
    import torch
    torch.cuda.set_device(0)
    torch.zeros(1, device="cuda")
if stage >= "c":
    import torch.distributed as dist
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29533")
    dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", 0))
    t = torch.ones(4, device="cuda"); dist.all_reduce(t)
n, p = 50000, 500000
g = m.B200SnpLinAlg.synthetic(n, p, 2024)
mk, mt = C.c_double(0), C.c_double(0)
m._lib.check(m.load().ihtb_sweep_bench(g._h, 0, 2, 5, C.byref(mk), C.byref(mt)))
print("stage", stage, "ok", mk.value, flush=True)
