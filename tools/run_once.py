"""Run the predict kernel a few times on device-resident synthetic inputs (profiling target).
    python tools/run_once.py [rows] [precision] [launches]
"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from oracle import refmath as rm  # noqa: E402  (synthetic inputs only)

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
prec = sys.argv[2] if len(sys.argv) > 2 else "bf16x3"
launches = int(sys.argv[3]) if len(sys.argv) > 3 else 3
emu_mod = importlib.import_module("21cmvae_b200.emulator")
pp = importlib.import_module("21cmvae_b200.preprocess")
kh = importlib.import_module("21cmvae_b200.keras_h5")
ks, bs, relu = rm.glorot_chain(rm.DIRECT_DIMS, seed=2022)
mu, sd = rm.synthetic_signal_stats(ks, bs, relu)
pmin, pmax = rm.prior_par_stats()
emu = emu_mod.DirectEmulator(stats=pp.NormStats(pmin, pmax, mu, sd))
emu.emulator = emu_mod.DenseModel(kh.DenseChainWeights(ks, bs, relu, name="emulator"))
params = torch.from_numpy(rm.draw_params(rows, seed=1)).cuda()
out = torch.empty((rows, 451), dtype=torch.float32, device="cuda")
if os.environ.get("VAE21_RUN_NORMALISED"):  # bypass the parameter transform: already-normalised float32 inputs straight into the Dense chain
    xn = torch.rand((rows, 7), dtype=torch.float32, device="cuda") * 2 - 1
    _p = emu.emulator.predict
    emu.predict = lambda _params, out=None, precision=None: _p(xn, out=out, precision=precision)
for _ in range(launches):
    emu.predict(params, out=out, precision=prec)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
emu.predict(params, out=out, precision=prec)
e1.record()
torch.cuda.synchronize()
print(f"{prec} rows={rows} ms={e0.elapsed_time(e1):.4f} checksum={float(out[::997].sum()):.4f}")
