"""Event timeline of one CTA pair of the tensor-core kernel, from a profiling build (-DVAE21_TC_TIMING=1):
    VAE21_LIB=tools/ab/libvae21_timing.so python tools/tc_trace.py [tiles_per_pair] [precision] [tile_to_print]
Roles: M = first MMA warp (leader CTA), E0 / E19 = first / last epilogue warp of the leader, F0 = first epilogue warp of the follower,
P = weight producer (leader).  Times are SM cycles relative to the first event of the printed tile."""
import ctypes as C
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from oracle import refmath as rm  # noqa: E402

tiles = int(sys.argv[1]) if len(sys.argv) > 1 else 8
prec = sys.argv[2] if len(sys.argv) > 2 else "fp16e4m3"
show = int(sys.argv[3]) if len(sys.argv) > 3 else 5
rows = 74 * 256 * tiles
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
n = lib.vae21_debug_tc_trace_len()
buf = (C.c_ulonglong * (5 * n))()
for _ in range(2):
    emu.predict(p, out=o, precision=prec)
torch.cuda.synchronize()
lib.vae21_debug_tc_trace(buf, 1)  # discard the warm-up launches
emu.predict(p, out=o, precision=prec)
torch.cuda.synchronize()
assert lib.vae21_debug_tc_trace(buf, 0) == 0
a = np.array(buf[:], dtype=np.uint64).reshape(5, n)
names = ["M", "E0", "P", "F0", "E19"]
tagname = {1: "chunk start", 2: "acc/operand ok", 3: "ring ok  k=", 4: "operands ok k=", 5: "issued nk=", 6: "chunk committed", 10: "wait chunk", 11: "woke chunk",
           12: "processed", 13: "signalled", 20: "wait empty slot", 21: "got slot"}
ev = []
for r in range(5):
    cnt = int(a[r, 0])
    for i in range(1, cnt + 1):
        e = int(a[r, i])
        ev.append((e & 0xffffffffff, r, e >> 56, (e >> 40) & 0xffff))
ev.sort()
# tile boundaries: M's "chunk start" with arg 0
starts = [t for t, r, tag, arg in ev if r == 0 and tag == 1 and arg == 0]
print(f"{prec}: {len(starts)} tiles traced on CTA pair 0; tile durations (cycles): {[starts[i + 1] - starts[i] for i in range(len(starts) - 1)]}")
if show + 1 < len(starts):
    t0, t1 = starts[show], starts[show + 1]
    # per-chunk summary from the MMA warp and epilogue warp E0
    print(f"--- tile {show}: {t1 - t0} cycles")
    last = {}
    for t, r, tag, arg in ev:
        if t < t0 - 2000 or t > t1 + 6000:
            continue
        if r == 2:
            continue  # the producer is printed separately
        key = (r,)
        dt = t - last.get(key, t)
        last[key] = t
        print(f"{t - t0:8d}  {names[r]:3s} +{dt:6d}  {tagname.get(tag, tag)} {arg}")
    # producer: time spent waiting for empty slots within the tile
    w = 0
    tw = None
    for t, r, tag, arg in ev:
        if r != 2 or t < t0 or t > t1:
            continue
        if tag == 20:
            tw = t
        elif tag == 21 and tw is not None:
            w += t - tw
            tw = None
    print(f"producer: {w} cycles of the tile waiting for empty ring slots")
