"""BASELINE config 4: emcee-style ensemble MCMC (stretch move, a = 2, red/blue halves) with 1e5 walkers; the per-step cost is two
batched emulator + likelihood evaluations (fused chi^2 kernel, nothing but 4 B/row leaves the GPU).  Walkers shard over the ranks of
one node (each rank's sub-ensemble is an independent stretch-move ensemble; no per-step communication).

    python tools/mcmc_bench.py [walkers=100000] [steps=1000] [precision=fp16e4m3]
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/mcmc_bench.py

Prints one JSON line: ms per MCMC step (the whole run as one library call, device-timed, max over ranks; and one call per step with a
host synchronisation), walker-updates/s, acceptance fraction, and the time of the two emulate + chi^2 launches of a step alone.
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
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
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

mc = importlib.import_module("21cmvae_b200.mcmc")
dev = torch.device("cuda", local)
# walkers live in the transformed prior box (log10 for fstar, Vc, fx); start in a small ball around the mid-point
s = mc.StretchMoveSampler(emu, obs, 25.0, pmin, pmax, walkers=W, seed=4, precision=prec, rank=rank, world=world)
n = s.n
s.ball((pmin + pmax) / 2, 0.05 * (pmax - pmin))
s.run(20)  # burn-in / warm-up (also computes the initial ln p)
torch.cuda.synchronize()
launches0 = h.info()["kernel_launches"]
if world > 1:
    dist.barrier()
# (a) the whole run as ONE library call, device-timed
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
e0.record()
frac = s.run(steps)
e1.record()
torch.cuda.synchronize()
run_ms = e0.elapsed_time(e1) / steps
# (b) one call per step with a host synchronisation (what a driver that inspects the chain every step pays)
times = []
for _ in range(min(steps, 100)):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    s.run(1)
    torch.cuda.synchronize()
    times.append(time.perf_counter() - t0)
times = np.array(times)
# the library's emulate + chi^2 kernel alone: two launches per step
x_norm = torch.rand((n // 2, 7), dtype=torch.float32, device=dev) * 2 - 1
chi = torch.empty(n // 2, dtype=torch.float32, device=dev)
phys = torch.from_numpy(rm.draw_params(n // 2, seed=1)).to(dev)
stream = torch.cuda.current_stream(dev)
torch.cuda.synchronize()
e0.record()
for _ in range(50):
    h.chi2(phys, obs, isig, out=chi, want_best=False, precision=P, stream=stream.cuda_stream)
e1.record()
torch.cuda.synchronize()
kernel_ms = e0.elapsed_time(e1) / 50 * 2
med = float(np.median(times))
mean, cov = s.moments()
if world > 1:
    t = torch.tensor([run_ms, med], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    run_ms, med = float(t[0]), float(t[1])
if rank == 0:
    print(json.dumps({"workload": "stretch-move ensemble MCMC, 7 parameters, flat prior box, fused chi^2 vs a 451-bin mock observation; "
                                  "the move runs in the library (vae21_mcmc_run: proposal kernel + fused emulate/chi^2 launch + accept kernel per half-step)",
                      "walkers": W, "n_gpus": world, "walkers_per_gpu": n, "steps": steps, "precision": prec,
                      "ms_per_step_one_call": run_ms, "ms_per_step_call_per_step_median": med * 1e3, "walker_updates_per_s": W / (run_ms * 1e-3),
                      "chi2_kernel_ms_per_step": kernel_ms, "kernel_launches_per_step": 6, "acceptance_fraction": frac,
                      "mean_chi2": float(-2 * s.logp.mean().item()), "posterior_mean": [float(v) for v in mean]}))
if world > 1:
    dist.destroy_process_group()
