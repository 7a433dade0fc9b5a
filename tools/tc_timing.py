"""Print the MMA-warp wait breakdown of a profiling build (-DVAE21_TC_TIMING=1):
    VAE21_LIB=tools/ab/libvae21_timing.so python tools/tc_timing.py [rows]"""
import ctypes as C
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from oracle import refmath as rm  # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
emu_mod = importlib.import_module("21cmvae_b200.emulator")
pp = importlib.import_module("21cmvae_b200.preprocess")
kh = importlib.import_module("21cmvae_b200.keras_h5")
L = importlib.import_module("21cmvae_b200._lib")
ks, bs, relu = rm.glorot_chain(rm.DIRECT_DIMS, seed=2022)
mu, sd = rm.synthetic_signal_stats(ks, bs, relu)
pmin, pmax = rm.prior_par_stats()
emu = emu_mod.DirectEmulator(stats=pp.NormStats(pmin, pmax, mu, sd))
emu.emulator = emu_mod.DenseModel(kh.DenseChainWeights(ks, bs, relu, name="emulator"))
p = torch.from_numpy(rm.draw_params(rows, seed=1)).cuda()
o = torch.empty((rows, 451), dtype=torch.float32, device="cuda")
for _ in range(3):
    emu.predict(p, out=o, precision=os.environ.get("VAE21_TIMING_PRECISION", "fp16e4m3"))
torch.cuda.synchronize()
buf = (C.c_longlong * (160 * 16))()
lib = L.load()
assert lib.vae21_debug_tc_timing(buf) == 0
a = np.array(buf[:], dtype=np.int64).reshape(160, 16)
a = a[a[:, 0] > 0]
names = ["total", "a0 wait", "accumulator-free wait", "ring wait", "operand wait", "issue blocks", "loop iterations (incl. waits)"]
m = a.mean(axis=0)
print(f"CTAs reporting: {len(a)}; mean cycles of the MMA warp per launch")
ntile = rows / 128 / 148
for i, n in enumerate(names):
    print(f"  {n:32s} {m[i]:12.0f}  {100 * m[i] / m[0]:5.1f}%   {m[i] / ntile:8.0f} cycles/tile")
