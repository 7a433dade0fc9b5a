"""Ensemble MCMC on top of the fused emulate + chi^2 kernel (BASELINE config 4).

The reference has no sampler: its users call ``DirectEmulator.predict`` (emulator.py:383-407) inside their own
likelihood, typically from emcee.  ``StretchMoveSampler`` is that loop moved onto the GPU -- positions, ln p, the
stretch move and the likelihood all stay device-resident (``vae21_mcmc_run``: per half-step one proposal kernel, one
fused emulate + chi^2 launch, one accept kernel), a run of K steps is one library call.

Walkers live in the coordinates of ``preprocess.par_transform``'s box (log10 on the masked columns), the likelihood is
Gaussian: ln p = -chi^2 / 2 inside the prior box, -inf outside.  Across the GPUs of a node each rank runs an independent
sub-ensemble (no per-step communication); posterior moments are combined with ``multigpu.allreduce_sums``.
"""

from __future__ import annotations

import numpy as np

from . import _lib
from . import multigpu


class StretchMoveSampler:
    def __init__(self, emulator, observed, sigma, lo, hi, walkers: int, seed: int = 0, a: float = 2.0, precision=None,
                 rank: int = 0, world: int = 1, group=None):
        import torch

        self.emulator = emulator
        self.handle = emulator._handle()
        self.nd = self.handle.dims[0]
        self.obs = np.ascontiguousarray(observed, dtype=np.float32)
        nout = self.handle.dims[-1]
        self.isig = np.ascontiguousarray(1.0 / np.broadcast_to(np.asarray(sigma, dtype=np.float64), (nout,))).astype(np.float32)
        self.lo = np.ascontiguousarray(np.broadcast_to(np.asarray(lo, np.float64), (self.nd,)), dtype=np.float64)
        self.hi = np.ascontiguousarray(np.broadcast_to(np.asarray(hi, np.float64), (self.nd,)), dtype=np.float64)
        self.a, self.seed = float(a), int(seed) + 7919 * int(rank)
        self.rank, self.world, self.group = int(rank), int(world), group
        lo_w, hi_w = multigpu.shard_bounds(int(walkers), self.world, self.rank)
        self.n = (hi_w - lo_w) // 2 * 2  # even sub-ensemble
        if self.n < 2:
            raise ValueError("need at least two walkers per rank")
        from .emulator import _resolve_precision

        self.precision = _resolve_precision(precision if precision is not None else emulator.precision)
        self.device = torch.device("cuda", self.handle.device)
        self.x = torch.empty((self.n, self.nd), dtype=torch.float64, device=self.device)
        self.logp = torch.empty((self.n,), dtype=torch.float64, device=self.device)
        self.step = 0
        self.accepted = 0
        self._need_init = True

    def set_positions(self, x):
        """(walkers_of_this_rank, n_par) start positions (host or device), coordinates of the prior box."""
        import torch

        t = torch.as_tensor(np.asarray(x, dtype=np.float64) if not hasattr(x, "device") else x, dtype=torch.float64).to(self.device)
        if tuple(t.shape) != (self.n, self.nd):
            raise ValueError(f"positions must have shape ({self.n}, {self.nd})")
        self.x.copy_(t)
        self._need_init = True

    def ball(self, centre, scale):
        """Start in a small Gaussian ball (emcee's usual initialisation), clipped to the box."""
        rng = np.random.default_rng(self.seed)
        x = np.asarray(centre, np.float64) + np.asarray(scale, np.float64) * rng.standard_normal((self.n, self.nd))
        self.set_positions(np.clip(x, self.lo, self.hi))

    def run(self, n_steps: int, sync: bool = True):
        """Advance the ensemble by n_steps (in place).  Returns the acceptance fraction of this call (None if not sync)."""
        import torch

        acc = self.handle.mcmc_run(self.x, self.logp, self.lo, self.hi, self.obs, self.isig, a=self.a, seed=self.seed,
                                   first_step=self.step, n_steps=int(n_steps), init_logp=self._need_init, precision=self.precision,
                                   stream=torch.cuda.current_stream(self.device).cuda_stream, want_accepted=sync)
        self._need_init = False
        self.step += int(n_steps)
        if acc is None:
            return None
        self.accepted += acc
        return acc / max(1, self.n * int(n_steps))

    def moments(self):
        """Posterior mean and covariance over ALL ranks' current walkers (one small all-reduce)."""
        x = self.x.double()
        s = np.concatenate([[float(x.shape[0])], x.sum(dim=0).cpu().numpy(), (x.T @ x).cpu().numpy().ravel()])
        s = multigpu.allreduce_sums(s, group=self.group)
        n, d = s[0], self.nd
        mean = s[1:1 + d] / n
        cov = s[1 + d:].reshape(d, d) / n - np.outer(mean, mean)
        return mean, cov
