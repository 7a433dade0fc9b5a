"""CPU: the N>1 host logic over gloo with world_size 2 (row sharding + global argmin)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, pkg


def test_shard_bounds_cover_exactly():
    mg = pkg("multigpu")
    for n in (0, 1, 7, 1000, 10**6 + 3):
        for world in (1, 2, 3, 8):
            blocks = [mg.shard_bounds(n, world, r) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        mg.shard_bounds(10, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import sys

    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import importlib

        mg = importlib.import_module("21cmvae_b200.multigpu")
        from oracle import refmath as rm

        ks, bs, relu = rm.glorot_chain((7, 16, 451), seed=3)
        mu, sd = rm.synthetic_signal_stats(ks, bs, relu)
        pmin, pmax = rm.prior_par_stats()
        params = rm.draw_params(1001, seed=8)
        truth = rm.predict(params[777], ks, bs, relu, pmin, pmax, mu, sd)
        sigma = np.full(451, 2.0)

        def cpu_chi2(p):  # oracle stands in for the CUDA chi2 in this CPU test
            return rm.chi2(rm.predict(p, ks, bs, relu, pmin, pmax, mu, sd, squeeze=False), truth, 1 / sigma)

        sh = mg.ShardedEmulator(None, rank, world)
        lo, hi = sh.local_block(len(params))
        bv, bi = sh.chi2_argmin(params, truth, sigma, compute_chi2=cpu_chi2)
        sums = mg.allreduce_sums(np.array([hi - lo, float(rank)]))
        # a rank with only NaNs must never win
        nv, ni = mg.global_argmin(float("nan") if rank == 0 else 3.5, -1 if rank == 0 else 4, lo)
        q.put((rank, lo, hi, bv, bi, sums.tolist(), nv, ni))
    finally:
        dist.destroy_process_group()


def test_world2_gloo_argmin_and_sums():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, lo0, hi0, bv0, bi0, s0, nv0, ni0), (r1, lo1, hi1, bv1, bi1, s1, nv1, ni1) = res
    assert (lo0, hi0, lo1, hi1) == (0, 501, 501, 1001)
    assert bi0 == bi1 == 777 and bv0 == bv1 and bv0 < 1e-12
    assert s0 == s1 == [1001.0, 1.0]
    assert nv0 == nv1 == 3.5 and ni0 == ni1 == 501 + 4
