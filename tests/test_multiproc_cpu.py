"""World-size-2 gloo tests (CPU) of the multi-GPU host logic: shard partition, CV farm and the SNP-sharded protocol."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT

WORLD = 2


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    return port


def _worker(rank, port, fn_name, ret):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    try:
        ret[rank] = globals()[fn_name](rank)
    finally:
        dist.destroy_process_group()


def _run(fn_name):
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(_free_port(), fn_name, ret), nprocs=WORLD, join=True)
        return dict(ret)


def _data(d, seed=31, n=600, p=900, k=4):
    from mendeliht_jl_b200 import synth
    y, z, *_ = synth.simulate_response(seed, n, p, k, d, n_cov=1)
    bed = synth.packed_columns(seed, n, np.arange(p))
    return y, z, bed, n, p, k


# ---- workers ----------------------------------------------------------------------------------------------------
def w_sharded_fit(rank):
    from mendeliht_jl_b200 import parallel
    from oracle import glm, iht, snp
    import sharded_sim
    out = {}
    for d, l in ((glm.NORMAL, glm.IDENTITY), (glm.BERNOULLI, glm.LOGIT)):
        y, z, bed, n, p, k = _data(d)
        j0, pl = parallel.shard_range(p, WORLD, rank)
        x_loc = snp.SnpLinAlgOracle(bed[j0:j0 + pl], n)
        v, best, it, tr = sharded_sim.fit_sharded(y, x_loc, j0, p, z, k + 1, d, l)
        ref = iht.fit_iht(y, snp.SnpLinAlgOracle(bed, n), z, k=k + 1, d=d, l=l)
        out[d] = (it == ref.iter, bool(np.array_equal(np.flatnonzero(v.best_b), np.flatnonzero(ref.beta))),
                  float(np.max(np.abs(v.best_b - ref.beta))), float(abs(best - ref.logl)),
                  tr.backtracks == ref.trace.backtracks)
    return out


def w_sharded_group_fit(rank):
    from mendeliht_jl_b200 import parallel
    from oracle import glm, iht, snp
    import sharded_sim
    y, z, bed, n, p, k = _data(glm.NORMAL, seed=17, n=500, p=800, k=4)
    group = np.arange(p) // 90 + 1                   # 9 groups; one of them straddles the shard boundary at 400
    j0, pl = parallel.shard_range(p, WORLD, rank)
    x_loc = snp.SnpLinAlgOracle(bed[j0:j0 + pl], n)
    v, best, it, tr = sharded_sim.fit_sharded(y, x_loc, j0, p, z, 2, glm.NORMAL, glm.IDENTITY, J=3, group=group)
    ref = iht.fit_iht(y, snp.SnpLinAlgOracle(bed, n), z, k=2, J=3, group=group)
    return (it == ref.iter, bool(np.array_equal(np.flatnonzero(v.best_b), np.flatnonzero(ref.beta))),
            float(np.max(np.abs(v.best_b - ref.beta))), int(np.count_nonzero(ref.beta)))


def w_cv_farm(rank):
    from mendeliht_jl_b200 import parallel, api
    from oracle import cv as ocv, glm, snp
    y, z, bed, n, p, k = _data(glm.NORMAL, seed=5, n=400, p=500, k=3)
    x = snp.SnpLinAlgOracle(bed, n)
    folds = 1 + (np.arange(n) % 3)
    path = [1, 2, 4]
    grid = api.allocate_fold_and_k(3, path)

    def cv_fn(combos):
        # stand-in for api.cv_iht(..., combos=combos): the same contract, computed by the oracle on this rank's share
        mses = np.zeros(len(grid)); iters = np.zeros(len(grid), dtype=np.int64)
        _, full_m, full_i = ocv.cv_iht(y, x, z, path=path, q=3, folds=folds, return_grid=True)
        for i in combos:
            mses[i], iters[i] = full_m[i], full_i[i]
        return mses, iters

    mses, iters = parallel.cv_iht_farm(dist, cv_fn, len(grid))
    ref_mse, ref_grid, ref_it = ocv.cv_iht(y, x, z, path=path, q=3, folds=folds, return_grid=True)
    mine = parallel.deal_round_robin(len(grid), WORLD, rank)
    return (bool(np.allclose(mses, ref_grid, rtol=0, atol=0)), bool(np.array_equal(iters, ref_it)),
            bool(np.allclose(api.meanloss(mses, 3, folds), ref_mse)), mine)


# ---- tests ------------------------------------------------------------------------------------------------------
def test_shard_range_partitions_columns():
    from mendeliht_jl_b200 import parallel
    for p, w in ((10, 3), (500000, 8), (7, 8), (1000001, 4)):
        cover = []
        for r in range(w):
            j0, pl = parallel.shard_range(p, w, r)
            cover += list(range(j0, j0 + pl)) if p < 100 else [(j0, pl)]
        if p < 100:
            assert cover == list(range(p))
        else:
            assert cover[0][0] == 0 and sum(c[1] for c in cover) == p
            assert all(cover[i][0] + cover[i][1] == cover[i + 1][0] for i in range(w - 1))
            assert max(c[1] for c in cover) - min(c[1] for c in cover) <= 1
    assert parallel.deal_round_robin(7, 3, 1) == [1, 4]


def test_sharded_protocol_matches_oracle_world2():
    res = _run("w_sharded_fit")
    for rank in range(WORLD):
        for d, (same_it, same_supp, dbeta, dlogl, same_bt) in res[rank].items():
            assert same_it and same_supp and same_bt, (rank, d)
            assert dbeta < 1e-10 and dlogl < 1e-8, (rank, d, dbeta, dlogl)


def test_cv_farm_world2():
    res = _run("w_cv_farm")
    assert res[0][3] == [0, 2, 4, 6, 8] and res[1][3] == [1, 3, 5, 7]
    for rank in range(WORLD):
        assert res[rank][0] and res[rank][1] and res[rank][2]


def test_sharded_group_protocol_matches_oracle_world2():
    """Doubly sparse projection over two shards (groups straddling the boundary): gathered per-group bounds and
    candidate lists reproduce the single-process oracle."""
    res = _run("w_sharded_group_fit")
    for rank in range(WORLD):
        same_it, same_supp, dbeta, nnz = res[rank]
        assert same_it and same_supp and dbeta < 1e-10 and 0 < nnz <= 6, (rank, res[rank])
