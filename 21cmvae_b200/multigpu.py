"""Row sharding of a parameter batch across the GPUs of one node (one process per GPU).

The emulator evaluates independent rows, so ``predict`` needs no collective: rank r of G takes
the contiguous block ``[r*N/G, (r+1)*N/G)`` with replicated weights and constants.  Collectives
(``torch.distributed``: NCCL on GPUs, gloo in the CPU tests) are used only for the tiny
reductions callers make on top: the global chi^2 argmin and posterior sums.
(The reference is single-process)
"""

from __future__ import annotations

from typing import Optional, Tuple

import numpy as np


def shard_bounds(n: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous, balanced block of rank ``rank``: sizes differ by at most one row."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, rem = divmod(int(n), world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def bind_host_to_gpu(device_index: int) -> Optional[list]:
    """Pin this process to the CPU cores NVML reports as local to GPU ``device_index`` (same NUMA node / PCIe
    root), intersected with the cores the container allows.  Pinned host buffers allocated afterwards land on
    that node, which matters for the PCIe-bound host<->device path when several ranks copy at once.
    Returns the core list used, or None when NVML / affinity information is unavailable (no-op)."""
    import os

    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(int(device_index))
        ncpu = os.cpu_count() or 64
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cores = [64 * w + b for w, m in enumerate(words) for b in range(64) if (m >> b) & 1]
        allowed = os.sched_getaffinity(0)
        use = sorted(set(cores) & set(allowed))
        if use:
            os.sched_setaffinity(0, use)
            return use
    except Exception:  # noqa: BLE001 - affinity is an optimisation only
        pass
    return None


def global_argmin(local_val: float, local_idx: int, row_offset: int, group=None) -> Tuple[float, int]:
    """Combine per-rank ``(min chi2, local row)`` into the global minimum and GLOBAL row index.
    One all_gather of 16 bytes per rank.  NaN / idx < 0 (no finite value on a rank) never wins."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return float(local_val), (int(local_idx) + row_offset if local_idx >= 0 else -1)
    world = dist.get_world_size(group)
    backend = dist.get_backend(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    gidx = float(local_idx + row_offset) if local_idx >= 0 else -1.0
    mine = torch.tensor([float(local_val), gidx], dtype=torch.float64, device=dev)
    allv = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(allv, mine, group=group)
    best_v, best_i = float("nan"), -1
    for t in allv:
        v, i = float(t[0]), int(t[1])
        if i < 0 or np.isnan(v):
            continue
        if best_i < 0 or v < best_v or (v == best_v and i < best_i):
            best_v, best_i = v, i
    return best_v, best_i


def allreduce_sums(values: np.ndarray, group=None) -> np.ndarray:
    """Sum a small float64 vector (posterior moments: weight, weighted means, second moments) over ranks."""
    import torch
    import torch.distributed as dist

    v = np.asarray(values, dtype=np.float64)
    if not (dist.is_available() and dist.is_initialized()):
        return v.copy()
    backend = dist.get_backend(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    t = torch.from_numpy(v.copy()).to(dev)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t.cpu().numpy()


class ShardedEmulator:
    """Evaluate this rank's block of a global parameter batch with a (Direct|AutoEncoder)Emulator.

    ``compute_chi2`` lets tests inject a CPU stand-in; the product path always uses the emulator's
    CUDA ``chi2``.
    """

    def __init__(self, emulator, rank: int = 0, world: int = 1, group=None):
        self.emulator = emulator
        self.rank, self.world, self.group = int(rank), int(world), group

    def local_block(self, n: int) -> Tuple[int, int]:
        return shard_bounds(n, self.world, self.rank)

    def predict_local(self, params_global, precision=None):
        lo, hi = self.local_block(len(params_global))
        return self.emulator.predict(params_global[lo:hi], precision=precision), (lo, hi)

    def chi2_argmin(self, params_global, observed, sigma, precision=None, compute_chi2=None):
        """Global (chi2_min, row) over all ranks; each rank evaluates only its block."""
        lo, hi = self.local_block(len(params_global))
        if hi > lo:
            if compute_chi2 is not None:
                c = np.asarray(compute_chi2(params_global[lo:hi]))
                ok = ~np.isnan(c)
                if ok.any():
                    i = int(np.nanargmin(c))
                    bv, bi = float(c[i]), i
                else:
                    bv, bi = float("nan"), -1
            else:
                _, bv, bi = self.emulator.chi2(params_global[lo:hi], observed, sigma, precision=precision,
                                               return_argmin=True)
        else:
            bv, bi = float("nan"), -1
        return global_argmin(bv, bi, lo, group=self.group)
