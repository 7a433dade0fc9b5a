"""Error of the tensor-core operand formats on TRAINED weights, simulated in float64 on the CPU (no GPU needed).

    python tools/precision_study.py [n_inputs=20000] [--json]

Networks: (1) the reference's shipped autoencoder-based emulator chain 7-352-352-352-224-9-32-352-451 (tests/golden/ae_chain.npz,
real trained weights) and (2) the DirectEmulator-shaped student trained on it (tests/golden/direct_trained.h5,
tools/make_trained_fixture.py).  Inputs: uniform in [-1, 1]^7 (the normalised prior box), seed 20211.

Formats (each layer: exact products of the rounded operands, float64 accumulation -- the fp32 accumulation of the hardware adds
~1e-7 relative, see the FP32 row):
  fp32      float32 operands and accumulation (the FP32-SIMT path)
  bf16x1 / fp16x1   one-pass 16-bit operands (shown to fail the budget: why the split formats exist)
  bf16x3 / fp16x3   a_hi w_hi + a_hi w_lo + a_lo w_hi with 16-bit hi / lo parts (tc_kernel.cuh FMT 0 / 1)
  fp16e4m3  a_hi w_hi [fp16] + e4m3(a) e4m3(w_lo S) + e4m3(a_lo 2^11) e4m3(w S / 2^11), accumulator at scale S = 2^11 (FMT 2)
Errors are against the float64 chain on the same (float32-rounded) weights and inputs, in sigma units (units of
np.std(signal_train)) and in mK for sigma = 50 mK resp. the fixture's own sigma; budget 0.01 mK rms / 0.05 mK max.
This is test infrastructure (the GPU tests assert the same budget on the real kernels: tests/test_gpu_parity.py).
"""
import importlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def rn_bf16(x):
    u = np.asarray(x, np.float32).view(np.uint32).astype(np.uint64)
    r = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16) << 16
    return r.astype(np.uint32).view(np.float32).astype(np.float64)


def rn_fp16(x):
    with np.errstate(over="ignore"):
        return np.asarray(x, np.float32).astype(np.float16).astype(np.float64)


def rn_e4m3(x):
    """Round to nearest-even e4m3 (bias 7, 3 mantissa bits, subnormals down to 2^-9), saturating at +-448 (satfinite)."""
    x = np.asarray(x, np.float64)
    a = np.abs(x)
    e = np.floor(np.log2(np.maximum(a, 2.0 ** -20)))
    e = np.maximum(e, -6.0)                 # subnormal range shares the exponent of the smallest normal
    q = 2.0 ** (e - 3)                      # spacing
    r = np.round(a / q) * q                 # np.round is round-half-even
    r = np.minimum(r, 448.0)
    return np.sign(x) * r


def chain(x, ks, bs, relu, fmt):
    h = np.asarray(x, np.float64)
    for k, b, r in zip(ks, bs, relu):
        w = np.asarray(k, np.float64)
        if fmt == "fp64":
            y = h @ w
        elif fmt == "fp32":
            y = (h.astype(np.float32) @ k.astype(np.float32)).astype(np.float64)
        elif fmt in ("bf16x1", "fp16x1"):
            rn = rn_bf16 if fmt[0] == "b" else rn_fp16
            y = rn(h) @ rn(w)
        elif fmt in ("bf16x3", "fp16x3"):
            rn = rn_bf16 if fmt[0] == "b" else rn_fp16
            ah, wh = rn(h), rn(w)
            al, wl = rn(h - ah), rn(w - wh)
            y = ah @ wh + ah @ wl + al @ wh
        elif fmt == "fp16e4m3":
            S = 2048.0
            while S > 1 and np.abs(w).max() * S > 32768.0:
                S *= 0.5
            ws = w * S
            wh = rn_fp16(ws)
            wl8, w8 = rn_e4m3(ws - wh), rn_e4m3(ws / 2048.0)
            ah = rn_fp16(h)
            a8, al8 = rn_e4m3(h), rn_e4m3((h - ah) * 2048.0)
            y = (ah @ wh + a8 @ wl8 + al8 @ w8) / S
        else:
            raise ValueError(fmt)
        h = y + np.asarray(b, np.float64)
        if r:
            h = np.maximum(h, 0.0)
    return h


def study(name, ks, bs, relu, n, sigma_mk):
    rng = np.random.default_rng(20211)
    x = rng.uniform(-1, 1, size=(n, ks[0].shape[0])).astype(np.float32)
    ref = chain(x, ks, bs, relu, "fp64")
    amp = np.max(np.abs(ref), axis=1)
    rows = []
    for fmt in ("fp32", "bf16x1", "fp16x1", "bf16x3", "fp16x3", "fp16e4m3"):
        d = chain(x, ks, bs, relu, fmt) - ref
        rms = np.sqrt(np.mean(d * d, axis=1))
        rows.append({"net": name, "format": fmt, "rms_mean_sigma": float(rms.mean()), "rms_max_sigma": float(rms.max()),
                     "max_abs_sigma": float(np.abs(d).max()), "max_over_amplitude": float((np.abs(d).max(axis=1) / amp).max()),
                     "rms_max_mK": float(rms.max() * sigma_mk), "max_abs_mK": float(np.abs(d).max() * sigma_mk),
                     "inside_budget": bool(rms.max() * sigma_mk <= 0.01 and np.abs(d).max() * sigma_mk <= 0.05)})
    pre = np.abs(ref).max()
    return rows, float(pre)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 20000
    kh = importlib.import_module("21cmvae_b200.keras_h5")
    g = np.load(os.path.join(ROOT, "tests", "golden", "ae_chain.npz"))
    nl = sum(1 for k in g.files if k.startswith("k") and k[1:].isdigit())
    out = []
    rows, _ = study("ae_chain (reference's trained weights)", [g[f"k{i}"] for i in range(nl)], [g[f"b{i}"] for i in range(nl)],
                    [bool(r) for r in g["relu"]], n, 50.0)
    out += rows
    w = kh.load_dense_chain(os.path.join(ROOT, "tests", "golden", "direct_trained.h5"))
    st = np.load(os.path.join(ROOT, "tests", "golden", "direct_trained.npz"))
    rows, _ = study("direct_trained (student of the AE chain)", w.kernels, w.biases, w.relu, n, float(st["sig_std"]))
    out += rows
    if "--json" in sys.argv:
        print(json.dumps({"n_inputs": n, "rows": out}))
        return
    print(f"{n} uniform inputs; errors vs the float64 chain; budget 0.01 mK rms / 0.05 mK max")
    print(f"{'network':44s} {'format':9s} {'rms mean (s)':>12s} {'rms max (s)':>12s} {'max abs (s)':>12s} {'max/amp':>10s} {'rms max mK':>11s} {'max mK':>9s}  ok")
    for r in out:
        print(f"{r['net']:44s} {r['format']:9s} {r['rms_mean_sigma']:12.2e} {r['rms_max_sigma']:12.2e} {r['max_abs_sigma']:12.2e} "
              f"{r['max_over_amplitude']:10.2e} {r['rms_max_mK']:11.2e} {r['max_abs_mK']:9.2e}  {'yes' if r['inside_budget'] else 'NO'}")


if __name__ == "__main__":
    main()
