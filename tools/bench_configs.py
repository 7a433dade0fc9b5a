"""Secondary measurements for BASELINE.json configs 3 and 4 (not the bench.py headline):
  * fused chi^2 + argmin throughput on device-resident parameter grids (no spectra written),
  * per-call latency of small batches (MCMC step: 1e5 walkers; single-signal call).
    python tools/bench_configs.py [precision]
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

from oracle import refmath as rm  # noqa: E402  (synthetic inputs only)

prec = sys.argv[1] if len(sys.argv) > 1 else "bf16x3"
emu_mod = importlib.import_module("21cmvae_b200.emulator")
pp = importlib.import_module("21cmvae_b200.preprocess")
kh = importlib.import_module("21cmvae_b200.keras_h5")
L = importlib.import_module("21cmvae_b200._lib")
ks, bs, relu = rm.glorot_chain(rm.DIRECT_DIMS, seed=2022)
mu, sd = rm.synthetic_signal_stats(ks, bs, relu)
pmin, pmax = rm.prior_par_stats()
emu = emu_mod.DirectEmulator(stats=pp.NormStats(pmin, pmax, mu, sd))
emu.emulator = emu_mod.DenseModel(kh.DenseChainWeights(ks, bs, relu, name="emulator"))
h = emu._handle()
P = L.PRECISIONS[prec]
truth = rm.predict(np.array([0.0003, 4.2, 0, 0.055, 1.0, 0.1, 10]), ks, bs, relu, pmin, pmax, mu, sd)
obs = (truth + np.random.default_rng(7).normal(size=451) * 25).astype(np.float32)
isig = np.full(451, 1 / 25.0, np.float32)
out = {"precision": prec}


def timed(fn, reps):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for n in (1_000_000, 10_000_000):
    p = torch.from_numpy(rm.draw_params(n, seed=3)).cuda()
    c = torch.empty(n, dtype=torch.float32, device="cuda")
    ms = timed(lambda: h.chi2(p, obs, isig, out=c, want_best=False, precision=P), 5)
    out[f"chi2_device_{n}"] = {"ms": ms, "signals_per_s": n / ms * 1e3}
    t0 = time.perf_counter()
    _, bv, bi = h.chi2(p, obs, isig, want_chi2=False, want_best=True, precision=P)
    out[f"chi2_argmin_only_{n}"] = {"ms_wall": (time.perf_counter() - t0) * 1e3, "best": bv, "row": bi}
    del p, c

for n in (100_000, 1024, 1):
    ph = L.pinned_empty((n, 7), np.float64)
    ph[:] = rm.draw_params(n, seed=4)
    pd = torch.from_numpy(np.ascontiguousarray(ph)).cuda()
    od = torch.empty((n, 451), dtype=torch.float32, device="cuda")
    cd = torch.empty(n, dtype=torch.float32, device="cuda")
    out[f"predict_device_{n}"] = {"ms": timed(lambda: h.predict(pd, out=od, precision=P), 20)}
    out[f"chi2_device_{n}"] = {"ms": timed(lambda: h.chi2(pd, obs, isig, out=cd, want_best=False, precision=P), 20)}
    t = []
    for _ in range(20):
        t0 = time.perf_counter()
        emu.predict(ph, precision=prec)
        t.append(time.perf_counter() - t0)
    out[f"predict_host_api_{n}"] = {"ms_median": float(np.median(t)) * 1e3}
    t = []
    for _ in range(20):
        t0 = time.perf_counter()
        emu.chi2(ph, obs, 25.0, precision=prec)
        t.append(time.perf_counter() - t0)
    out[f"chi2_host_api_{n}"] = {"ms_median": float(np.median(t)) * 1e3}
print(json.dumps(out, indent=1))
