"""Training loop behind ``DirectEmulator.train`` (reference: VeryAccurateEmulator/emulator.py:339-381, which calls
``tf.keras.Model.fit(batch_size=256, validation_batch_size=256)``; optimiser and callbacks as in
notebooks/Training.ipynb cells 4-5).  The arithmetic runs in the library's CUDA kernels (csrc/train_kernels.cuh) through the
C-ABI trainer; this file is the host-side schedule Keras provides in the reference: shuffling, batching, the epoch loop,
validation, callbacks, and -- for data-parallel retraining (BASELINE config 5) -- the gradient all-reduce over the ranks of one
node (torch.distributed: NCCL on GPUs).

Keras semantics kept: every epoch visits a fresh permutation in batches of ``batch_size`` (the last one may be short); the
step minimises the MEAN of the per-sample losses of the batch; the reported training loss of an epoch is the sample-weighted
mean of the batch losses seen during the epoch; validation runs after each epoch; ``callbacks`` see ``on_epoch_end`` with
``{"loss", "val_loss", "lr"}`` and may stop training or change the learning rate.

Data parallelism: each global batch of ``batch_size`` rows is split into contiguous shares, one per rank; every rank
scales its gradient by 1 / (n_out * global batch rows) so that ONE sum all-reduce yields the gradient Keras would compute on
a single device, and all ranks apply the same Adam update (parameters stay bit-identical across ranks).
"""
from __future__ import annotations

import math
import os
import time
from typing import List, Optional, Sequence

import numpy as np

from . import _lib


class Adam:
    """Optimiser settings with the tf.keras.optimizers.Adam defaults of Keras 2.x."""

    def __init__(self, learning_rate=0.001, beta_1=0.9, beta_2=0.999, epsilon=1e-7):
        self.learning_rate = float(np.float32(learning_rate))  # a float32 variable in Keras
        self.beta_1 = float(beta_1)
        self.beta_2 = float(beta_2)
        self.epsilon = float(epsilon)
        self.iterations = 0
        # slot variables (flat, parameter order) of the last fit: the next fit of the same model continues from them, as a
        # Keras optimiser object does; keras_h5 stores / restores them (`optimizer_weights`)
        self.m = None
        self.v = None

    @classmethod
    def from_state(cls, st):
        """From a ``keras_h5.AdamState`` (what ``load_model`` finds in a saved model)."""
        opt = cls(st.learning_rate, st.beta_1, st.beta_2, st.epsilon)
        opt.iterations = int(st.iterations)
        opt.m = None if st.m is None else np.array(st.m, np.float32)
        opt.v = None if st.v is None else np.array(st.v, np.float32)
        return opt


class Callback:
    def on_train_begin(self, state):
        pass

    def on_epoch_end(self, epoch, logs, state):
        pass

    def on_train_end(self, state):
        pass


class EarlyStopping(Callback):
    """tf.keras.callbacks.EarlyStopping for a quantity that should decrease (``monitor='val_loss'``)."""

    def __init__(self, monitor="val_loss", min_delta=0.0, patience=0, restore_best_weights=False, verbose=0):
        self.monitor, self.min_delta, self.patience = monitor, abs(float(min_delta)), int(patience)
        self.restore_best_weights, self.verbose = bool(restore_best_weights), verbose
        self.best, self.wait, self.best_weights, self.stopped_epoch = math.inf, 0, None, None

    def on_train_begin(self, state):
        self.best, self.wait, self.best_weights, self.stopped_epoch = math.inf, 0, None, None

    def on_epoch_end(self, epoch, logs, state):
        cur = logs.get(self.monitor)
        if cur is None:
            return
        if self.restore_best_weights and self.best_weights is None:
            self.best_weights = state.get_params()
        self.wait += 1
        if cur < self.best - self.min_delta:
            self.best, self.wait = cur, 0
            if self.restore_best_weights:
                self.best_weights = state.get_params()
        if self.wait >= self.patience and epoch > 0:
            self.stopped_epoch = epoch
            state.stop_training = True

    def on_train_end(self, state):
        if self.stopped_epoch is not None and self.restore_best_weights and self.best_weights is not None:
            state.set_params(self.best_weights)
        if self.stopped_epoch is not None and self.verbose:
            print(f"Epoch {self.stopped_epoch + 1}: early stopping")


class ReduceLROnPlateau(Callback):
    """tf.keras.callbacks.ReduceLROnPlateau (mode 'min')."""

    def __init__(self, monitor="val_loss", factor=0.1, patience=10, verbose=0, min_delta=1e-4, cooldown=0, min_lr=0.0):
        if factor >= 1.0:
            raise ValueError("ReduceLROnPlateau does not support a factor >= 1.0.")
        self.monitor, self.factor, self.patience, self.verbose = monitor, float(factor), int(patience), verbose
        self.min_delta, self.cooldown, self.min_lr = float(min_delta), int(cooldown), float(min_lr)
        self.best, self.wait, self.cooldown_counter = math.inf, 0, 0

    def on_train_begin(self, state):
        self.best, self.wait, self.cooldown_counter = math.inf, 0, 0

    def on_epoch_end(self, epoch, logs, state):
        cur = logs.get(self.monitor)
        if cur is None:
            return
        if self.cooldown_counter > 0:
            self.cooldown_counter -= 1
            self.wait = 0
        if cur < self.best - self.min_delta:
            self.best, self.wait = cur, 0
        elif self.cooldown_counter <= 0:
            self.wait += 1
            if self.wait >= self.patience:
                old = state.optimizer.learning_rate
                if old > np.float32(self.min_lr):
                    new = max(float(np.float32(old * self.factor)), self.min_lr)
                    state.optimizer.learning_rate = new
                    if self.verbose:
                        print(f"Epoch {epoch + 1}: ReduceLROnPlateau reducing learning rate to {new}.")
                    self.cooldown_counter, self.wait = self.cooldown, 0


class FitState:
    """What callbacks may touch: the optimiser, the stop flag and the parameters."""

    def __init__(self, trainer, optimizer):
        self.trainer, self.optimizer, self.stop_training = trainer, optimizer, False

    def get_params(self):
        return self.trainer.get_params()

    def set_params(self, flat):
        self.trainer.set_params(flat, reset_moments=False)


def shard_batch(lo: int, hi: int, world: int, rank: int):
    """Contiguous share [a, b) of the batch positions [lo, hi) for this rank (earlier ranks take the remainder)."""
    n = hi - lo
    base, rem = divmod(n, world)
    a = lo + rank * base + min(rank, rem)
    return a, a + base + (1 if rank < rem else 0)


def fit(dims: Sequence[int], relu: Sequence[int], flat_params: np.ndarray, x, y, w, x_val=None, y_val=None, w_val=None, *,
        optimizer: Optional[Adam] = None, epochs=1, batch_size=256, callbacks: Sequence[Callback] = (), shuffle=True, seed=None,
        device=0, distributed=False, verbose=0):
    """Train a Dense stack.  x (n, n_in), y (n, n_out): already normalised / preprocessed float32; w (n,): per-sample loss weights
    1 / amplitude^2.  Returns (flat parameters, history dict with 'loss', 'val_loss', 'lr')."""
    import torch

    optimizer = optimizer or Adam()
    dev = torch.device("cuda", device)
    world, rank = 1, 0
    dist = None
    if distributed:
        import torch.distributed as dist  # noqa: WPS433

        if not dist.is_initialized():
            raise RuntimeError("distributed=True needs an initialised torch.distributed process group")
        world, rank = dist.get_world_size(), dist.get_rank()
    tr = _lib.Trainer(dims, relu, max_batch=batch_size, device=device)
    tr.set_params(flat_params)
    if optimizer.m is not None and optimizer.v is not None:
        if np.size(optimizer.m) != tr.num_params or np.size(optimizer.v) != tr.num_params:
            raise ValueError("the optimizer's moments belong to another model (parameter count differs)")
        tr.set_moments(optimizer.m, optimizer.v)
    n, n_out = int(np.shape(x)[0]), int(dims[-1])
    as_dev = lambda a: torch.as_tensor(np.ascontiguousarray(a, dtype=np.float32)).to(dev)  # noqa: E731
    dx, dy, dw = as_dev(x), as_dev(y), as_dev(w)
    has_val = x_val is not None and len(x_val) > 0
    if has_val:
        vx, vy, vw = as_dev(x_val), as_dev(y_val), as_dev(w_val)
    grad = torch.zeros(tr.num_params, dtype=torch.float32, device=dev)
    loss_acc = torch.zeros(1, dtype=torch.float32, device=dev)
    stream = torch.cuda.current_stream(dev).cuda_stream
    rng = np.random.default_rng(seed)
    state = FitState(tr, optimizer)
    hist = {"loss": [], "val_loss": [], "lr": []}
    timing = {"allreduce_s": 0.0} if os.environ.get("VAE21_TRAIN_TIMING") else None
    # experiment / test switches: VAE21_TRAIN_PER_BATCH=1 runs the data-parallel per-batch schedule on one GPU as well (no collective);
    # VAE21_TRAIN_DP_GRAPH=1 replays that schedule's step as two captured graphs around the all-reduce instead of issuing its 23
    # kernels one by one.  Same bits; measured SLOWER on 2 x B200 (0.0372 vs 0.0277 s per epoch, profiles/r2_train_n2_*.json: the
    # step is bound by the GPU's small-kernel latency, not by the host's launch rate), hence opt-in
    force_per_batch = bool(os.environ.get("VAE21_TRAIN_PER_BATCH"))
    use_graphs = bool(os.environ.get("VAE21_TRAIN_DP_GRAPH"))
    for cb in callbacks:
        cb.on_train_begin(state)
    for epoch in range(int(epochs)):
        perm = rng.permutation(n) if shuffle else np.arange(n)
        if distributed:  # every rank must walk the same permutation
            pt = torch.as_tensor(perm, dtype=torch.int64, device=dev)
            dist.broadcast(pt, src=0)
            perm = pt.cpu().numpy()
        d_perm = torch.as_tensor(perm.astype(np.int32), device=dev)
        loss_acc.zero_()
        per_batch = world > 1 or force_per_batch
        if not per_batch:
            # one GPU: the whole epoch in one library call (no Python between batches)
            tr.epoch(dx, dy, dw, d_perm, n, batch_size, optimizer.learning_rate, optimizer.beta_1, optimizer.beta_2, optimizer.epsilon,
                     optimizer.iterations, loss_acc, stream=stream)
            optimizer.iterations += (n + batch_size - 1) // batch_size
        n_graph = 0
        if per_batch and use_graphs and n >= 2 * batch_size:
            # data parallel: every FULL batch is two graph replays around the all-reduce (this rank's share of a batch is the same
            # position range in every batch), the learning rates of the epoch's updates are precomputed on the device
            a0, b0 = shard_batch(0, batch_size, world, rank)
            tr.dp_begin(dx, dy, dw, d_perm, n, batch_size, a0, b0 - a0, optimizer.learning_rate, optimizer.beta_1, optimizer.beta_2,
                        optimizer.epsilon, optimizer.iterations, grad, loss_acc, stream=stream)
            n_graph = n // batch_size
            for _ in range(n_graph):
                if b0 > a0:
                    tr.dp_forward_backward(stream=stream)
                else:
                    grad.zero_()
                if distributed and world > 1:
                    if timing is not None:
                        torch.cuda.synchronize()
                        t_a = time.perf_counter()
                    dist.all_reduce(grad, op=dist.ReduceOp.SUM)
                    if timing is not None:
                        torch.cuda.synchronize()
                        timing["allreduce_s"] += time.perf_counter() - t_a
                tr.dp_adam(stream=stream)
            optimizer.iterations += n_graph
        for lo in (range(n_graph * batch_size, n, batch_size) if per_batch else ()):
            hi = min(lo + batch_size, n)
            a, b = shard_batch(lo, hi, world, rank)
            if b > a:
                tr.forward_backward(dx, dy, dw, b - a, 1.0 / (n_out * (hi - lo)), grad, loss_acc, idx=d_perm[a:b], stream=stream)
            else:
                grad.zero_()
            if distributed and world > 1:
                if timing is not None:
                    torch.cuda.synchronize()
                    t_a = time.perf_counter()
                dist.all_reduce(grad, op=dist.ReduceOp.SUM)
                if timing is not None:
                    torch.cuda.synchronize()
                    timing["allreduce_s"] += time.perf_counter() - t_a
            optimizer.iterations += 1
            t = optimizer.iterations
            # (the betas reach the kernels as C floats; the library's epoch / dp_begin use those values here too)
            b1f, b2f = float(np.float32(optimizer.beta_1)), float(np.float32(optimizer.beta_2))
            lr_t = float(np.float32(optimizer.learning_rate)) * math.sqrt(1.0 - b2f**t) / (1.0 - b1f**t)
            tr.adam(grad, lr_t, optimizer.beta_1, optimizer.beta_2, optimizer.epsilon, stream=stream)
        if distributed and world > 1:
            dist.all_reduce(loss_acc, op=dist.ReduceOp.SUM)
        logs = {"loss": float(loss_acc.item()) / n, "lr": optimizer.learning_rate}
        if has_val:
            loss_acc.zero_()
            nv = int(vx.shape[0])
            for lo in range(0, nv, batch_size):
                hi = min(lo + batch_size, nv)
                tr.forward_backward(vx, vy, vw, hi - lo, 0.0, None, loss_acc, first=lo, stream=stream)
            logs["val_loss"] = float(loss_acc.item()) / nv
        hist["loss"].append(logs["loss"])
        hist["lr"].append(logs["lr"])
        if has_val:
            hist["val_loss"].append(logs["val_loss"])
        if verbose:
            print(f"Epoch {epoch + 1}/{epochs} - loss: {logs['loss']:.4e}" + (f" - val_loss: {logs['val_loss']:.4e}" if has_val else ""))
        for cb in callbacks:
            cb.on_epoch_end(epoch, logs, state)
        if state.stop_training:
            break
    for cb in callbacks:
        cb.on_train_end(state)
    out = tr.get_params()
    optimizer.m, optimizer.v = tr.get_moments()
    hist["kernel_launches"] = tr.launches()
    if timing is not None:
        hist["timing"] = timing
    tr.close()
    return out, hist


def flatten_weights(kernels: List[np.ndarray], biases: List[np.ndarray]) -> np.ndarray:
    return np.concatenate([np.concatenate([np.asarray(k, np.float32).ravel(), np.asarray(b, np.float32).ravel()])
                           for k, b in zip(kernels, biases)])


def unflatten_weights(flat: np.ndarray, dims: Sequence[int]):
    ks, bs, off = [], [], 0
    for l in range(len(dims) - 1):
        cnt = dims[l] * dims[l + 1]
        ks.append(np.array(flat[off:off + cnt], np.float32).reshape(dims[l], dims[l + 1]))
        off += cnt
        bs.append(np.array(flat[off:off + dims[l + 1]], np.float32))
        off += dims[l + 1]
    return ks, bs
