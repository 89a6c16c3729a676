"""Multi-GPU host logic, one process per GPU (launched by torchrun / torch.distributed.run).

Two patterns from BASELINE.json:
  * SNP-sharded fit (configs[4], no reference equivalent): columns block-partitioned over the ranks; the C library
    all-reduces the partial X*beta n-vectors and all-gathers top-k candidates over its own NCCL communicator, whose
    128-byte id is broadcast here through torch.distributed (plumbing only).
  * cross-validation farm (configs[2]): every rank holds the full matrix; the q x |path| grid of independent fits
    (reference `Threads.@threads :static` loop, src/cross_validation.jl:98-121) is dealt round-robin and the
    per-(fold,k) losses are summed across ranks.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _lib
from ._lib import check, load


def shard_range(p: int, world: int, rank: int):
    """Contiguous block partition of p columns: (j0, p_local); sizes differ by at most one."""
    base, rem = divmod(int(p), int(world))
    j0 = rank * base + min(rank, rem)
    return j0, base + (1 if rank < rem else 0)


def deal_round_robin(n_items: int, world: int, rank: int):
    """Grid positions a rank runs in the CV farm."""
    return list(range(rank, n_items, world))


def nccl_library_path():
    """The NCCL that torch bundles (so the library and torch share one libnccl in the process)."""
    try:
        import nvidia.nccl
        for base in list(getattr(nvidia.nccl, "__path__", [])):
            cand = os.path.join(base, "lib", "libnccl.so.2")
            if os.path.exists(cand):
                return cand
    except Exception:
        pass
    return None


class Comm:
    """ihtb_comm handle.  `dist` is torch.distributed with an initialised process group."""

    def __init__(self, dist, device_index: int):
        import torch
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        lib = load()
        check(lib.ihtb_set_device(device_index))
        path = nccl_library_path()
        cpath = path.encode() if path else None
        uid = np.zeros(128, dtype=np.uint8)
        if self.rank == 0:
            check(lib.ihtb_comm_unique_id(cpath, uid.ctypes.data_as(C.POINTER(C.c_uint8))))
        backend = dist.get_backend()
        t = torch.from_numpy(uid)
        if backend == "nccl":
            t = t.cuda(device_index)
        dist.broadcast(t, src=0)
        uid = t.cpu().numpy().astype(np.uint8)
        self._h = C.c_void_p()
        check(lib.ihtb_comm_create(cpath, uid.ctypes.data_as(C.POINTER(C.c_uint8)), self.rank, self.world,
                                   C.byref(self._h)))

    def allreduce_latency_us(self, n: int, reps: int = 200, p2p: bool = True) -> float:
        """Collective micro-benchmark (device time per all-reduce of n doubles); every rank must call it."""
        out = C.c_double(0.0)
        check(load().ihtb_comm_allreduce_bench(self._h, n, reps, 1 if p2p else 0, C.byref(out)))
        return out.value

    def close(self):
        if getattr(self, "_h", None):
            load().ihtb_comm_destroy(self._h)
            self._h = None


def cv_iht_farm(dist, cv_fn, n_grid: int):
    """Run `cv_fn(combos) -> (mses[n_grid], iters[n_grid])` on this rank's share of the grid and sum over ranks.
    Positions a rank does not run must be left at 0."""
    import torch
    rank, world = dist.get_rank(), dist.get_world_size()
    combos = deal_round_robin(n_grid, world, rank)
    mses, iters = cv_fn(combos)
    buf = torch.from_numpy(np.concatenate([np.asarray(mses, dtype=np.float64),
                                           np.asarray(iters, dtype=np.float64)]))
    if dist.get_backend() == "nccl":
        buf = buf.cuda()
    dist.all_reduce(buf)
    out = buf.cpu().numpy()
    return out[:n_grid], out[n_grid:].astype(np.int64)


def bench_sharded(args, rank, world, local_rank, G):
    """bench.py at N > 1: weak scaling, P_PER_GPU columns per rank, SNP-sharded fit_iht."""
    import json
    import time
    import torch
    import torch.distributed as dist
    from . import api, synth

    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    lib = load()
    check(lib.ihtb_set_device(local_rank))
    comm = Comm(dist, local_rank)
    n, k = G["N_SAMPLES"], G["K_SPARSITY"]
    p = G["P_PER_GPU"] * world
    j0, p_local = shard_range(p, world, rank)
    g = api.B200SnpLinAlg.synthetic(n, p_local, G["SEED"], 0.0, j0)      # generated on the device, shard by shard
    y, z, true_idx, _, _ = synth.simulate_response(G["SEED"] + 1, n, p, k, G["DIST"], geno_seed=G["SEED"])
    clocks = G["ClockSampler"](local_rank)

    v = api.IHTVariable(g, z, y, k, G["DIST"], G["LINK"], comm=comm, p_global=p)
    for _ in range(args.warmup):
        v.init_iht_indices(None)
        v.fit(trace_cap=0)
    dist.barrier(); torch.cuda.synchronize()
    if rank == 0:
        clocks.start()
    l0 = _lib.launch_count()
    ms = C.c_double(0.0)
    check(lib.ihtb_fit_timer(v._h, 0, None))
    iters = sweeps = 0
    sweep_s = 0.0
    for _ in range(args.steps):
        v.init_iht_indices(None)
        res, _ = v.fit(trace_cap=0)
        iters += int(res.iter); sweeps += int(res.n_sweeps); sweep_s += res.sweep_seconds
    check(lib.ihtb_fit_timer(v._h, 1, C.byref(ms)))
    dist.barrier(); torch.cuda.synchronize()
    launches = _lib.launch_count() - l0
    ph = (C.c_double * 4)()
    check(lib.ihtb_fit_phase_times(v._h, ph))
    n_it_all = max(iters * (args.steps + args.warmup) // max(args.steps, 1), 1)
    phases = {k: ph[i] / n_it_all * 1e3 for i, k in enumerate(["stepsize_ms", "gradstep_ms", "xb_glm_ms", "score_sweep_ms"])}
    t = torch.tensor([ms.value], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    t_value = float(t.item()) * 1e-3
    beta, c, _, _ = v.get()
    v.close()

    # e2e: public API with host y / z every step
    t0 = time.perf_counter()
    e_iters = 0
    for _ in range(args.steps):
        r = api.fit_iht(y, g, z, k=k, d=G["DIST"], l=G["LINK"], comm=comm, p_global=p)
        e_iters += r.iter
    dist.barrier(); torch.cuda.synchronize()
    te = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    dist.all_reduce(te, op=dist.ReduceOp.MAX)
    t_e2e = float(te.item())
    clk = clocks.stop() if rank == 0 else None

    mk, mt = C.c_double(0.0), C.c_double(0.0)
    check(lib.ihtb_sweep_bench(g._h, _lib.SWEEP_FAST, 3, 20, C.byref(mk), C.byref(mt)))
    tk = torch.tensor([mk.value], dtype=torch.float64, device="cuda")
    dist.all_reduce(tk, op=dist.ReduceOp.MAX)
    if rank == 0:
        peak, peak_src = G["measured_peak"]()
        abytes = G["sweep_bytes"](n, p_local)
        achieved = abytes / (float(tk.item()) * 1e-3) / 1e9
        nz = np.flatnonzero(beta)
        line = {
            # weak scaling: every rank sweeps its own 500k-SNP shard each iteration, so the job processes
            # world x iterations shard-iterations (at N=1 this is plain iterations/s)
            "metric": "iht_iterations_per_sec", "value": world * iters / t_value, "unit": G["UNIT"], "n_gpus": world,
            "global_iterations_per_sec": iters / t_value,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": t_value / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": G["workload_config"](world),
            "iterations_per_fit": iters / args.steps, "sweeps_per_fit": sweeps / args.steps,
            "sweep_ms_in_fit": sweep_s / max(sweeps - args.steps, 1) * 1e3, "sweep_share_of_step": sweep_s / t_value,
            "host_phase_ms_per_iteration": phases,
            "packed_bytes_swept_per_sec_all_gpus": sweeps * G["sweep_bytes"](n, p) / t_value,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": None, "peak_source": peak_src, "kernel": "k_sweep_lut (per GPU, slowest rank)",
                         "algorithmic_bytes_per_launch": abytes, "kernel_ms": float(tk.item())},
            "e2e": {"value": world * e_iters / t_e2e, "unit": G["UNIT"], "global_iterations_per_sec": e_iters / t_e2e,
                    "h2d_bytes_per_step": int(y.nbytes + z.nbytes),
                    "d2h_bytes_per_step": int(beta.nbytes + c.nbytes), "ms_per_step": t_e2e / args.steps * 1e3,
                    "note": "fit_iht(y, x_shard, z; comm) on every rank with host y/z, global beta copied back; "
                            "genotype shards generated on the device (host generation of N x 6.25 GB is skipped)"},
            "gpu_launches": int(launches), "clocks": clk, "cpu_baseline": None,
            "collectives": "NCCL allreduce(n doubles) per X*beta, allgather of top-k candidates, allreduce of "
                           "re-scored candidates",
            "check": {"support_size": int(nz.size), "true_positives": int(np.intersect1d(nz, true_idx).size),
                      "iterations": iters // args.steps},
        }
        print(json.dumps(line))
    comm.close()
    dist.destroy_process_group()
