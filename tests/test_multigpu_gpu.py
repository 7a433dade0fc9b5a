"""GPU, world size 2 over NCCL (skipped on boxes with one GPU): the collectives BASELINE configs 3 and 5 use -- the global chi^2
argmin of a sharded parameter batch, and the gradient all-reduce of data-parallel retraining (reference emulator.py:339-381)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from conftest import GOLDEN, ROOT, pkg

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("", 0))  # free on every interface: the TCPStore of rank 0 listens on all of them
    p = s.getsockname()[1]
    s.close()
    return p


def _need_two_gpus():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs on one node")


def _guarded(worker, rank, world, port, q):
    """Run a worker; a failure is reported through the queue at once (the parent must not sit out a long timeout)."""
    try:
        worker(rank, world, port, q)
    except BaseException as e:  # noqa: BLE001
        q.put((rank, ("__error__", f"{type(e).__name__}: {e}")))
        raise


def _spawn(worker, world=2, timeout=240):
    """Two ranks of `worker`; the rendezvous port is retried when somebody else grabbed it between the probe and the listen."""
    last = None
    for attempt in range(3):
        ctx = mp.get_context("spawn")
        q = ctx.Queue()
        port = _free_port()
        procs = [ctx.Process(target=_guarded, args=(worker, r, world, port, q)) for r in range(world)]
        for p in procs:
            p.start()
        res, err = {}, None
        try:
            while len(res) < world and err is None:
                r, payload = q.get(timeout=timeout)
                if isinstance(payload, tuple) and len(payload) == 2 and payload[0] == "__error__":
                    err = payload[1]
                else:
                    res[r] = payload
        finally:
            for p in procs:
                p.join(timeout=5 if err else 60)
                if p.is_alive():
                    p.kill()
        if err is None:
            assert all(p.exitcode == 0 for p in procs)
            return res
        last = err
        if "EADDRINUSE" not in err and "address already in use" not in err:
            break
    raise AssertionError(f"worker failed: {last}")


def _init(rank, world, port):
    import sys

    import torch.distributed as dist

    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    return dist


def _trained_emulator(device):
    import importlib

    emu = importlib.import_module("21cmvae_b200.emulator")
    pp = importlib.import_module("21cmvae_b200.preprocess")
    d = np.load(os.path.join(GOLDEN, "direct_trained.npz"))
    e = emu.DirectEmulator(stats=pp.NormStats(d["par_min"], d["par_max"], d["sig_mean"], np.float32(d["sig_std"])), device=device)
    e.load_model(os.path.join(GOLDEN, "direct_trained.h5"))
    return e, d


def _argmin_worker(rank, world, port, q):
    dist = _init(rank, world, port)
    try:
        import importlib

        mg = importlib.import_module("21cmvae_b200.multigpu")
        from oracle import refmath as rm

        e, d = _trained_emulator(rank)
        params = rm.draw_params(20_001, seed=77)
        params[15_432] = d["par_test"][5]
        obs = e.predict(d["par_test"][5], precision="fp32")
        out = {}
        for prec in ("fp32", "fp16e4m3"):
            sh = mg.ShardedEmulator(e, rank, world)
            out[prec] = sh.chi2_argmin(params, obs, np.full(451, 3.0), precision=prec)
        single = e.chi2(params, obs, np.full(451, 3.0), precision="fp32", return_argmin=True)[1:] if rank == 0 else None
        sums = mg.allreduce_sums(np.array([1.0, float(rank)]))
        q.put((rank, (out, single, sums.tolist())))
    finally:
        dist.destroy_process_group()


def test_nccl_global_argmin_of_a_sharded_batch():
    _need_two_gpus()
    res = _spawn(_argmin_worker)
    (o0, single, s0), (o1, _, s1) = res[0], res[1]
    assert s0 == s1 == [2.0, 1.0]
    for prec in ("fp32", "fp16e4m3"):
        assert o0[prec] == o1[prec]                      # every rank holds the same global answer
        assert o0[prec][1] == 15_432                     # the planted row, in GLOBAL numbering (it lives on rank 1)
    assert o0["fp32"][0] == pytest.approx(single[0], rel=1e-6, abs=1e-6) and single[1] == 15_432
    assert o0["fp32"][0] < 1e-3


def _dp_worker(rank, world, port, q):
    dist = _init(rank, world, port)
    try:
        import importlib

        tr = importlib.import_module("21cmvae_b200.training")
        from oracle import refmath as rm

        dims = (7, 288, 352, 288, 224, 451)
        ks, bs, relu = rm.glorot_chain(dims, seed=5)
        rng = np.random.default_rng(6)
        n = 700  # two full batches of 256 and a short one
        x = rng.uniform(-1, 1, size=(n, 7)).astype(np.float32)
        y = rng.normal(size=(n, 451)).astype(np.float32)
        w = (1.0 / rng.uniform(0.5, 2.0, size=n) ** 2).astype(np.float32)
        flat = tr.flatten_weights(ks, bs)
        out, hist = tr.fit(dims, [int(r) for r in relu], flat, x, y, w, optimizer=tr.Adam(1e-3), epochs=2, batch_size=256, seed=9,
                           device=rank, distributed=True)
        # the same run with every full batch replayed as two captured graphs around the NCCL all-reduce (opt-in schedule)
        os.environ["VAE21_TRAIN_DP_GRAPH"] = "1"
        try:
            out_g, _ = tr.fit(dims, [int(r) for r in relu], flat, x, y, w, optimizer=tr.Adam(1e-3), epochs=2, batch_size=256, seed=9,
                              device=rank, distributed=True)
        finally:
            del os.environ["VAE21_TRAIN_DP_GRAPH"]
        ref = None
        if rank == 0:
            ref, hist1 = tr.fit(dims, [int(r) for r in relu], flat, x, y, w, optimizer=tr.Adam(1e-3), epochs=2, batch_size=256, seed=9,
                                device=0, distributed=False)
            ref = (ref, hist1["loss"])
        q.put((rank, (out, hist["loss"], ref, out_g)))
    finally:
        dist.destroy_process_group()


def test_nccl_data_parallel_training_equals_single_gpu():
    _need_two_gpus()
    res = _spawn(_dp_worker)
    (p0, l0, ref, g0), (p1, l1, _, g1) = res[0], res[1]
    assert np.array_equal(p0, p1)  # replicas stay bit-identical: same all-reduced gradient, same Adam
    assert np.array_equal(g0, p0) and np.array_equal(g1, p1)  # graph replay around the all-reduce: the same bits
    ref_p, ref_l = ref
    scale = np.abs(ref_p).max()
    assert np.max(np.abs(p0 - ref_p)) <= 2e-4 * scale  # sharded batch sums differ from the full-batch sum only in rounding
    assert np.allclose(l0, ref_l, rtol=1e-4) and np.allclose(l0, l1, rtol=1e-6)


def _grid_worker(rank, world, port, q):
    dist = _init(rank, world, port)
    try:
        import importlib

        mg = importlib.import_module("21cmvae_b200.multigpu")

        e, d = _trained_emulator(rank)
        npd = 6  # 6^7 = 279,936 grid nodes, generated in the kernel prologue
        total = npd**7
        obs = e.predict(d["par_test"][9], precision="fp32")
        lo, hi = mg.shard_bounds(total, world, rank)
        out = {}
        for prec in ("fp32", "bf16x3"):
            bv, bi, _ = e.chi2_grid(npd, obs, 5.0, first=lo, count=hi - lo, precision=prec)
            out[prec] = mg.global_argmin(bv, bi, 0)  # chi2_grid already returns GLOBAL grid indices
        whole = e.chi2_grid(npd, obs, 5.0, precision="fp32")[:2] if rank == 0 else None
        q.put((rank, (out, whole, (lo, hi))))
    finally:
        dist.destroy_process_group()


def test_nccl_argmin_over_a_device_generated_grid():
    """BASELINE config 3 in miniature: the grid is sharded by index range, every rank evaluates its range with the fused chi^2
    kernel on nodes generated on the device, the global minimum is one 16-byte all_gather."""
    _need_two_gpus()
    res = _spawn(_grid_worker)
    (o0, whole, b0), (o1, _, b1) = res[0], res[1]
    assert b0[1] == b1[0] and b0[0] == 0 and b1[1] == 6**7
    for prec in ("fp32", "bf16x3"):
        assert o0[prec] == o1[prec]
    assert o0["fp32"][1] == whole[1] and o0["fp32"][0] == pytest.approx(whole[0], rel=1e-6)
    assert o0["bf16x3"][1] == whole[1] and o0["bf16x3"][0] == pytest.approx(whole[0], rel=1e-3, abs=1e-3)
