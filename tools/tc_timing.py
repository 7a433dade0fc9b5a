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
lib = L.load()
rbuf = (C.c_ulonglong * (3 * 256))()
for i in range(3):
    if i == 2 and hasattr(lib, "vae21_debug_tc_rec_timing"):
        torch.cuda.synchronize()
        lib.vae21_debug_tc_rec_timing(rbuf, 1)  # discard the warm-up launches
    emu.predict(p, out=o, precision=os.environ.get("VAE21_TIMING_PRECISION", "fp16e4m3"))
torch.cuda.synchronize()
buf = (C.c_longlong * (160 * 16))()
assert lib.vae21_debug_tc_timing(buf) == 0
a = np.array(buf[:], dtype=np.int64).reshape(160, 16)
a = a[a[:, 0] > 0]
names = ["total", "operand / accumulator waits (after the baton)", "ring wait", "baton wait", "issue blocks"]
ntile = rows / 128 / 148
print(f"CTAs reporting: {len(a)} (leaders of the pairs); mean cycles per launch of the two MMA-issuing warps")
for w in range(2):
    m = a[:, 8 * w:8 * w + 8].mean(axis=0)
    if m[0] <= 0:
        continue
    print(f" issuer {w}")
    for i, n in enumerate(names):
        print(f"  {n:48s} {m[i]:12.0f}  {100 * m[i] / m[0]:5.1f}%   {m[i] / ntile:8.0f} cycles/tile")
    rest = m[0] - m[1:5].sum()
    print(f"  {'everything else (decode, loop, fences)':48s} {rest:12.0f}  {100 * rest / m[0]:5.1f}%   {rest / ntile:8.0f} cycles/tile")

print(f" prologue warp (leader CTAs; accumulated over the 3 launches): operand computation {a[:, 5].mean() / ntile / 3:8.0f} cycles/tile, "
      f"wait for a0 release {a[:, 6].mean() / ntile / 3:8.0f} cycles/tile")
if hasattr(lib, "vae21_debug_tc_rec_timing") and lib.vae21_debug_tc_rec_timing(rbuf, 0) == 0:
    r = np.array(rbuf[:], dtype=np.float64).reshape(3, 256)
    n = int((r[2] > 0).sum())
    print(f"per issue-table record (mean cycles per visit): {n} records per tile")
    print("  rec  flagged-wait  ring-wait")
    for i in range(n):
        v = max(r[2, i], 1)
        mark = " <--" if r[0, i] / v > 800 or r[1, i] / v > 400 else ""
        print(f"  {i:3d}  {r[0, i] / v:10.0f}  {r[1, i] / v:9.0f}{mark}")
