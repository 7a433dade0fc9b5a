"""Training (reference emulator.py:51-83, :339-381, notebooks/Training.ipynb cells 4-5).

CPU tests pin the oracle (oracle/train_ref.py) against torch autograd -- an independent derivative of the same loss -- and
check the host-side schedule (callbacks, batch sharding, the world-2 gradient all-reduce over gloo).  GPU tests compare the
CUDA trainer with the oracle step by step through the C-ABI.
"""
import math
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, pkg


def _problem(dims, n, seed, rm):
    ks, bs, relu = rm.glorot_chain(dims, seed=seed)
    rng = np.random.default_rng(seed + 1)
    bs = [rng.normal(scale=0.05, size=b.shape).astype(np.float32) for b in bs]  # non-zero biases exercise db
    x = rng.uniform(-1, 1, size=(n, dims[0])).astype(np.float32)
    y = rng.normal(size=(n, dims[-1])).astype(np.float32)
    mean_over_std = rng.normal(scale=0.3, size=dims[-1]).astype(np.float32)
    return ks, bs, relu, x, y, mean_over_std


# ---- oracle pinned against torch autograd ---------------------------------------------------------
def test_oracle_gradient_matches_torch_autograd(rm):
    from oracle import train_ref as tref

    dims = (7, 24, 40, 17)
    ks, bs, relu, x, y, mos = _problem(dims, 50, 5, rm)
    w = tref.sample_weights(y, mos)
    loss_rows, g = tref.loss_and_grad(x, y, w, ks, bs, relu, 1.0 / (dims[-1] * len(x)))
    tk = [torch.tensor(k, dtype=torch.float64, requires_grad=True) for k in ks]
    tb = [torch.tensor(b, dtype=torch.float64, requires_grad=True) for b in bs]
    h = torch.tensor(x, dtype=torch.float64)
    for k, b, r in zip(tk, tb, relu):
        h = h @ k + b
        if r:
            h = torch.relu(h)
    ty = torch.tensor(y, dtype=torch.float64)
    amp = torch.max(torch.abs(ty + torch.tensor(mos, dtype=torch.float64)), dim=1).values  # emulator.py:70-76
    per_sample = torch.mean((ty - h) ** 2, dim=1) / amp**2                                   # :78-80
    per_sample.mean().backward()                                                               # Keras minimises the batch mean
    want = np.concatenate([np.concatenate([k.grad.numpy().ravel(), b.grad.numpy().ravel()]) for k, b in zip(tk, tb)])
    assert np.allclose(loss_rows, per_sample.detach().numpy(), rtol=1e-12, atol=0)
    assert np.allclose(g, want, rtol=1e-10, atol=1e-14)
    # and the numpy loss of the API mirror agrees with it
    emu = pkg("emulator")
    sig_train = (y * 2.0 + 1.0)  # any training set: only its mean/std enter
    f = emu.relative_mse_loss(sig_train)
    m = np.mean(sig_train, axis=0) / np.std(sig_train)
    assert np.allclose(f(y, h.detach().numpy()), np.mean((y - h.detach().numpy()) ** 2, axis=1) / np.max(np.abs(y + m), axis=1) ** 2)


def test_oracle_adam_is_keras_adam():
    from oracle import train_ref as tref

    rng = np.random.default_rng(0)
    p, m, v = rng.normal(size=10), np.zeros(10), np.zeros(10)
    p0 = p.copy()
    g1, g2 = rng.normal(size=10), rng.normal(size=10)
    tref.adam_step(p, m, v, g1, 0.01, 1)
    # first step of Adam moves every weight by ~lr against the gradient sign
    assert np.allclose(p, p0 - 0.01 * g1 / (np.abs(g1) + 1e-7 / math.sqrt(1 - 0.999)), rtol=1e-9)
    tref.adam_step(p, m, v, g2, 0.01, 2)
    m2 = 0.9 * 0.1 * g1 + 0.1 * g2
    v2 = 0.999 * 0.001 * g1**2 + 0.001 * g2**2
    assert np.allclose(m, m2) and np.allclose(v, v2)


# ---- host-side schedule ------------------------------------------------------------------------------
def test_shard_batch_partitions_every_batch():
    tr = pkg("training")
    for lo, hi in ((0, 256), (256, 300), (0, 1), (10, 13)):
        for world in (1, 2, 3, 8):
            parts = [tr.shard_batch(lo, hi, world, r) for r in range(world)]
            assert parts[0][0] == lo and parts[-1][1] == hi
            assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1 and min(sizes) >= 0


class _FakeState:
    def __init__(self, lr):
        tr = pkg("training")
        self.optimizer = tr.Adam(lr)
        self.stop_training = False
        self.params = np.zeros(3, np.float32)

    def get_params(self):
        return self.params.copy()

    def set_params(self, p):
        self.params = np.array(p)


def test_reduce_lr_on_plateau_follows_keras():
    tr = pkg("training")
    cb = tr.ReduceLROnPlateau(monitor="val_loss", patience=2, factor=0.95, min_delta=5e-9, min_lr=1e-4)
    st = _FakeState(0.01)
    cb.on_train_begin(st)
    lrs = []
    for e, v in enumerate([1.0, 0.9, 0.9, 0.9, 0.9, 0.9, 0.8, 0.8, 0.8]):
        cb.on_epoch_end(e, {"val_loss": v}, st)
        lrs.append(st.optimizer.learning_rate)
    # improvement at epochs 0, 1, 6; two stagnant epochs -> reduce at epochs 3, 5 and 8
    assert lrs[2] == pytest.approx(0.01) and lrs[3] == pytest.approx(0.0095) and lrs[4] == pytest.approx(0.0095)
    assert lrs[5] == pytest.approx(0.0095 * 0.95) and lrs[7] == pytest.approx(0.0095 * 0.95) and lrs[8] == pytest.approx(0.0095 * 0.95**2)
    # the float32 storage of the Keras variable (notebooks/Training.ipynb prints 0.009499999787658453)
    assert lrs[3] == float(np.float32(0.01 * 0.95))
    with pytest.raises(ValueError):
        tr.ReduceLROnPlateau(factor=1.0)


def test_early_stopping_restores_best_weights():
    tr = pkg("training")
    cb = tr.EarlyStopping(monitor="val_loss", patience=2, min_delta=1e-10, restore_best_weights=True)
    st = _FakeState(0.01)
    cb.on_train_begin(st)
    stopped = None
    for e, v in enumerate([1.0, 0.5, 0.6, 0.7, 0.1]):
        st.params = np.full(3, float(e), np.float32)
        cb.on_epoch_end(e, {"val_loss": v}, st)
        if st.stop_training:
            stopped = e
            break
    cb.on_train_end(st)
    assert stopped == 3 and np.all(st.params == 1.0)  # best epoch was 1


# ---- world-2 gloo: sharded gradients sum to the single-device gradient ---------------------------------
def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _dp_worker(rank, world, port, q):
    import sys

    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import importlib

        tr = importlib.import_module("21cmvae_b200.training")
        from oracle import refmath as rm
        from oracle import train_ref as tref

        dims = (7, 12, 9)
        ks, bs, relu, x, y, mos = _problem(dims, 37, 11, rm)
        w = tref.sample_weights(y, mos)
        a, b = tr.shard_batch(0, 37, world, rank)
        _, g = tref.loss_and_grad(x[a:b], y[a:b], w[a:b], ks, bs, relu, 1.0 / (dims[-1] * 37))  # oracle stands in for CUDA
        t = torch.from_numpy(g.copy())
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        q.put((rank, t.numpy()))
    finally:
        dist.destroy_process_group()


def test_world2_gloo_sharded_gradient_equals_full_batch(rm):
    from oracle import train_ref as tref

    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_dp_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=180) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    dims = (7, 12, 9)
    ks, bs, relu, x, y, mos = _problem(dims, 37, 11, rm)
    w = tref.sample_weights(y, mos)
    _, full = tref.loss_and_grad(x, y, w, ks, bs, relu, 1.0 / (dims[-1] * 37))
    assert np.allclose(res[0], full, rtol=1e-12, atol=1e-15) and np.array_equal(res[0], res[1])


# ---- GPU: the CUDA trainer against the oracle ----------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("dims,batch", [((7, 288, 352, 288, 224, 451), 256), ((7, 33, 451), 77), ((5, 64, 64, 3), 1)])
def test_trainer_steps_match_oracle(rm, dims, batch):
    from oracle import train_ref as tref

    L = pkg("_lib")
    tr = pkg("training")
    n = batch * 3 + 5
    ks, bs, relu, x, y, mos = _problem(dims, n, 21, rm)
    w = tref.sample_weights(y, mos).astype(np.float32)
    t = L.Trainer(dims, relu, max_batch=256)
    flat0 = tr.flatten_weights(ks, bs)
    t.set_params(flat0)
    dev = torch.device("cuda", 0)
    dx, dy, dw = (torch.as_tensor(a).to(dev) for a in (x, y, w))
    grad = torch.zeros(t.num_params, dtype=torch.float32, device=dev)
    loss = torch.zeros(1, dtype=torch.float32, device=dev)
    p = flat0.astype(np.float64)
    m, v = np.zeros_like(p), np.zeros_like(p)
    rng = np.random.default_rng(3)
    for step in range(1, 5):
        idx = rng.permutation(n)[:batch].astype(np.int32)
        didx = torch.as_tensor(idx).to(dev)
        loss.zero_()
        t.forward_backward(dx, dy, dw, batch, 1.0 / (dims[-1] * batch), grad, loss, idx=didx)
        # gradient: oracle evaluated at the trainer's CURRENT parameters (Adam's first steps move every weight by ~lr times the
        # SIGN of its gradient, so two trajectories that differ by one rounding of a near-zero gradient separate by 2 lr)
        kso, bso = tref.unflatten(t.get_params().astype(np.float64), dims)
        rows, g = tref.loss_and_grad(x[idx], y[idx], w[idx], kso, bso, relu, 1.0 / (dims[-1] * batch))
        got_g = grad.cpu().numpy().astype(np.float64)
        scale = np.abs(g).max()
        assert np.abs(got_g - g).max() <= 2e-5 * scale, f"gradient off at step {step}"
        assert float(loss.item()) == pytest.approx(rows.sum(), rel=2e-5)
        # update: the oracle's Adam fed with the SAME (device) gradients must walk the same path
        lr_t = tref.adam_step(p, m, v, got_g, 0.01, step)
        t.adam(grad, lr_t)
        got_p = t.get_params().astype(np.float64)
        assert np.abs(got_p - p).max() <= 2e-5 * 0.01 * step + 1e-7, f"parameters drift at step {step}"
    # validation pass: forward + loss only, contiguous rows
    loss.zero_()
    t.forward_backward(dx, dy, dw, batch, 0.0, None, loss, first=2)
    kso, bso = tref.unflatten(t.get_params().astype(np.float64), dims)
    rows, _ = tref.loss_and_grad(x[2:2 + batch], y[2:2 + batch], w[2:2 + batch], kso, bso, relu, 0.0)
    assert float(loss.item()) == pytest.approx(rows.sum(), rel=2e-5)
    assert t.launches() > 0
    t.close()


@pytest.mark.gpu
def test_direct_emulator_train_end_to_end(rm):
    """emu.emulator.compile(...); emu.train(epochs, callbacks) like notebooks/Training.ipynb, on a synthetic data set generated
    by a teacher network; the loss must fall, callbacks must act, and predict must use the trained weights."""
    emu = pkg("emulator")
    tr = pkg("training")
    rng = np.random.default_rng(0)
    par = rm.draw_params(1500, seed=9, zero_fx_frac=0.0)
    tk, tb, trelu = rm.glorot_chain((7, 16, 451), seed=4)
    pmin, pmax = rm.prior_par_stats()
    sig = (rm.dense_chain(rm.par_transform_cached(par, pmin, pmax), tk, tb, trelu) * 40.0 - 60.0 + rng.normal(size=(1500, 451))).astype(np.float32)
    e = emu.DirectEmulator(par[:1200], par[1200:1400], par[1400:], sig[:1200], sig[1200:1400], sig[1400:], hidden_dims=[32, 32])
    before = e.test_error().mean()
    e.emulator.compile(optimizer=tr.Adam(0.01), loss=emu.relative_mse_loss(e.signal_train))
    cbs = [tr.EarlyStopping(monitor="val_loss", patience=15, min_delta=1e-10, restore_best_weights=True),
           tr.ReduceLROnPlateau(monitor="val_loss", patience=5, factor=0.95, min_delta=5e-9, min_lr=1e-4)]
    loss, val_loss = e.train(epochs=30, callbacks=cbs, verbose=0, seed=1)
    assert len(loss) == len(val_loss) == 30
    assert loss[-1] < 0.2 * loss[0] and val_loss[-1] < 0.2 * val_loss[0]
    assert e.test_error().mean() < 0.5 * before
    # the reported training loss is the relative MSE of the API's numpy loss (same formula, emulator.py:51-83)
    x = pkg("preprocess").par_transform(e.par_val, e.par_train)
    y = pkg("preprocess").preproc(e.signal_val, e.signal_train)
    pred = e.emulator.predict(x.astype(np.float32), precision="fp32")
    want = emu.relative_mse_loss(e.signal_train)(y, pred).mean()
    assert val_loss[-1] == pytest.approx(float(want), rel=1e-3)


@pytest.mark.gpu
def test_epoch_call_equals_per_batch_calls(rm):
    """vae21_trainer_epoch (all batches in one library call) is the same arithmetic as forward_backward + adam per batch."""
    L = pkg("_lib")
    tr = pkg("training")
    dims, batch, n = (7, 40, 24, 451), 64, 64 * 3 + 17
    ks, bs, relu, x, y, mos = _problem(dims, n, 33, rm)
    from oracle import train_ref as tref

    w = tref.sample_weights(y, mos).astype(np.float32)
    dev = torch.device("cuda", 0)
    dx, dy, dw = (torch.as_tensor(a).to(dev) for a in (x, y, w))
    perm = torch.as_tensor(np.random.default_rng(1).permutation(n).astype(np.int32)).to(dev)
    flat0 = tr.flatten_weights(ks, bs)
    a, b = L.Trainer(dims, relu, max_batch=batch), L.Trainer(dims, relu, max_batch=batch)
    a.set_params(flat0)
    b.set_params(flat0)
    grad = torch.zeros(a.num_params, dtype=torch.float32, device=dev)
    la, lb = torch.zeros(1, device=dev), torch.zeros(1, device=dev)
    it, lr = 0, float(np.float32(0.01))  # the epoch call takes the learning rate as a C float
    for lo in range(0, n, batch):
        rows = min(batch, n - lo)
        a.forward_backward(dx, dy, dw, rows, 1.0 / (dims[-1] * rows), grad, la, idx=perm[lo:lo + rows])
        it += 1
        b1, b2 = float(np.float32(0.9)), float(np.float32(0.999))  # ... and the betas as C floats
        a.adam(grad, lr * math.sqrt(1 - b2**it) / (1 - b1**it))
    b.epoch(dx, dy, dw, perm, n, batch, lr, 0.9, 0.999, 1e-7, 0, lb)
    assert np.array_equal(a.get_params(), b.get_params())
    assert float(la.item()) == float(lb.item())


@pytest.mark.gpu
def test_fit_continues_from_the_optimizers_saved_state(rm, tmp_path):
    """Two epochs in one fit == one epoch, save (weights + training_config + optimizer_weights), load, one more epoch: the Adam
    moments and the iteration count travel through the Keras HDF5 file, like tf.keras.Model.save / load_model + fit."""
    tr = pkg("training")
    kh = pkg("keras_h5")
    dims, n = (7, 40, 24, 451), 64 * 3 + 17
    ks, bs, relu, x, y, mos = _problem(dims, n, 5, rm)
    from oracle import train_ref as tref

    w = tref.sample_weights(y, mos).astype(np.float32)
    flat0 = tr.flatten_weights(ks, bs)
    one = tr.Adam(0.01)
    ref, _ = tr.fit(dims, relu, flat0, x, y, w, optimizer=one, epochs=2, batch_size=64, shuffle=False)
    opt = tr.Adam(0.01)
    half, _ = tr.fit(dims, relu, flat0, x, y, w, optimizer=opt, epochs=1, batch_size=64, shuffle=False)
    assert opt.iterations == 4 and opt.m is not None and np.any(opt.m != 0)
    k1, b1 = tr.unflatten_weights(half, dims)
    path = str(tmp_path / "half.h5")
    kh.save_dense_chain(path, kh.DenseChainWeights(k1, b1, [bool(r) for r in relu], name="emulator"), optimizer=opt)
    w2 = kh.load_dense_chain(path)
    opt2 = tr.Adam.from_state(kh.load_optimizer_state(path, w2))
    assert opt2.iterations == 4 and np.array_equal(opt2.m, opt.m) and np.array_equal(opt2.v, opt.v)
    done, _ = tr.fit(dims, relu, tr.flatten_weights(w2.kernels, w2.biases), x, y, w, optimizer=opt2, epochs=1, batch_size=64, shuffle=False)
    assert opt2.iterations == 8 == one.iterations
    assert np.array_equal(done, ref)
    assert np.array_equal(opt2.m, one.m) and np.array_equal(opt2.v, one.v)
    # and a fresh optimiser (moments reset) does NOT reproduce it: the state matters
    cold, _ = tr.fit(dims, relu, tr.flatten_weights(w2.kernels, w2.biases), x, y, w, optimizer=tr.Adam(0.01), epochs=1, batch_size=64, shuffle=False)
    assert not np.array_equal(cold, ref)


@pytest.mark.gpu
def test_data_parallel_graph_schedule_equals_the_other_schedules(rm, monkeypatch):
    """The data-parallel schedule (per batch: graph A = this rank's share forward/backward, [all-reduce], graph B = Adam) run on ONE
    GPU, with and without graph replay, gives the bits of the single-call epoch: three routes to the same arithmetic.  (Two ranks
    with NCCL between A and B: tests/test_multigpu_gpu.py.)"""
    tr = pkg("training")
    dims, n = (7, 40, 24, 451), 64 * 5 + 17
    ks, bs, relu, x, y, mos = _problem(dims, n, 11, rm)
    from oracle import train_ref as tref

    w = tref.sample_weights(y, mos).astype(np.float32)
    flat0 = tr.flatten_weights(ks, bs)
    kw = dict(epochs=3, batch_size=64, seed=3, x_val=x[:50], y_val=y[:50], w_val=w[:50])
    ref, h_ref = tr.fit(dims, relu, flat0, x, y, w, optimizer=tr.Adam(0.01), **kw)
    monkeypatch.setenv("VAE21_TRAIN_PER_BATCH", "1")
    plain, h_plain = tr.fit(dims, relu, flat0, x, y, w, optimizer=tr.Adam(0.01), **kw)
    monkeypatch.setenv("VAE21_TRAIN_DP_GRAPH", "1")
    graph, h_graph = tr.fit(dims, relu, flat0, x, y, w, optimizer=tr.Adam(0.01), **kw)
    assert np.array_equal(graph, plain), "graph replay differs from the same kernels launched one by one"
    assert np.array_equal(graph, ref), "per-batch schedule differs from the single-call epoch"
    assert h_graph["loss"] == h_plain["loss"] == h_ref["loss"] and h_graph["val_loss"] == h_ref["val_loss"]
    # the graph schedule runs the same kernels (counted per replay) plus its step counter
    assert h_graph["kernel_launches"] >= h_plain["kernel_launches"]
