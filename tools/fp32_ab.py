"""A/B of the two FP32 kernels (block-barrier `f32k` vs barrier-free `f32p`, csrc/fp32_pipe_kernel.cuh): the outputs must be
BIT-IDENTICAL (same products, same k order), only the time may differ.
    python tools/fp32_ab.py [name=path/to/libvae21_variant.so ...]
runs itself once per kernel (VAE21_FP32_PIPE=0 / 1, plus one child per extra library given as name=path: builds of the pipe
kernel with other compile-time switches, loaded through VAE21_LIB), compares every saved array, prints one JSON line.
The child mode (`--child TAG`) evaluates, for the DirectEmulator stack and the reference's trained AE chain: predict on ragged
row counts (f64 and f32 parameters), fused chi^2 + argmin, the fused error, and times 1M-row launches.
"""
import importlib
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child(tag, outdir, rows, launches):
    import torch

    from oracle import refmath as rm  # synthetic inputs only

    emu_mod = importlib.import_module("21cmvae_b200.emulator")
    pp = importlib.import_module("21cmvae_b200.preprocess")
    kh = importlib.import_module("21cmvae_b200.keras_h5")
    ks, bs, relu = rm.glorot_chain(rm.DIRECT_DIMS, seed=2022)
    mu, sd = rm.synthetic_signal_stats(ks, bs, relu)
    pmin, pmax = rm.prior_par_stats()
    emu = emu_mod.DirectEmulator(stats=pp.NormStats(pmin, pmax, mu, sd))
    emu.emulator = emu_mod.DenseModel(kh.DenseChainWeights(ks, bs, relu, name="emulator"))
    res = {}
    for n in (1, 7, 63, 64, 65, 300, 148 * 64 + 1, 100_003):
        p = rm.draw_params(n, seed=n)
        res[f"predict64_{n}"] = emu.predict(p, precision="fp32")
        res[f"predict32_{n}"] = emu.predict(p.astype(np.float32), precision="fp32")
    p = rm.draw_params(20_000, seed=5)
    truth = res["predict64_300"][17].astype(np.float64) + 0.25
    c, bv, bi = emu.chi2(p, truth, np.full(451, 25.0), precision="fp32", return_argmin=True)
    res["chi2"], res["chi2_best"] = c, np.array([bv, bi], np.float64)
    par = rm.draw_params(333, seed=41)
    emu.par_test, emu.signal_test = par, (emu.predict(par, precision="fp32") * 1.01 + 0.3).astype(np.float32)
    res["err_rel"] = emu.test_error(precision="fp32")
    res["err_abs_band"] = emu.test_error(precision="fp32", relative=False, flow=60.0, fhigh=120.0)
    # the reference's trained AE chain (8 layers, widths 352 ... 9 ... 451)
    g = np.load(os.path.join(ROOT, "tests", "golden", "ae_chain.npz"))
    nl = sum(1 for k in g.files if k.startswith("k") and k[1:].isdigit())
    ae = emu_mod.DenseModel(kh.DenseChainWeights([g[f"k{i}"] for i in range(nl)], [g[f"b{i}"] for i in range(nl)],
                                                 [bool(r) for r in g["relu"]], name="ae_chain"))
    res["ae_golden"] = ae.predict(g["x"], precision="fp32")
    xs = np.random.default_rng(3).uniform(-1, 1, size=(5001, 7)).astype(np.float32)
    res["ae_5001"] = ae.predict(xs, precision="fp32")
    # timing: device-resident 1M rows
    pd = torch.from_numpy(rm.draw_params(rows, seed=1)).cuda()
    od = torch.empty((rows, 451), dtype=torch.float32, device="cuda")
    for _ in range(2):
        emu.predict(pd, out=od, precision="fp32")
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(launches):
        emu.predict(pd, out=od, precision="fp32")
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / launches
    res["big_sample"] = od[::4099].cpu().numpy()
    res["big_checksum"] = np.array([float(od.double().sum())])
    np.savez(os.path.join(outdir, f"{tag}.npz"), **res)
    print(json.dumps({"tag": tag, "rows": rows, "ms_per_launch": ms, "tflops": rows * 740608 / ms / 1e9}))


def main():
    rows = int(os.environ.get("AB_ROWS", 1_000_000))
    launches = int(os.environ.get("AB_LAUNCHES", 5))
    if len(sys.argv) > 2 and sys.argv[1] == "--child":
        child(sys.argv[2], sys.argv[3], rows, launches)
        return
    outdir = tempfile.mkdtemp(prefix="fp32_ab_")
    lines = {}
    runs = [("barrier", "0", None), ("pipe", "1", None)]
    for spec in sys.argv[1:]:
        name, path = spec.split("=", 1)
        runs.append((name, "1", os.path.abspath(path)))
    for tag, val, lib in runs:
        env = dict(os.environ, VAE21_FP32_PIPE=val)
        if lib:
            env["VAE21_LIB"] = lib
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", tag, outdir], env=env, capture_output=True,
                           text=True, timeout=600)
        if r.returncode != 0:
            print(json.dumps({"tag": tag, "failed": r.returncode, "stderr": r.stderr[-2000:]}))
            sys.exit(1)
        lines[tag] = json.loads(r.stdout.strip().splitlines()[-1])
    a = np.load(os.path.join(outdir, "barrier.npz"))
    differing = []
    for tag, _, _ in runs[1:]:
        b = np.load(os.path.join(outdir, f"{tag}.npz"))
        differing += [f"{tag}:{k}" for k in a.files if not np.array_equal(a[k].view(np.uint8) if a[k].dtype.kind == "f" else a[k],
                                                                          b[k].view(np.uint8) if b[k].dtype.kind == "f" else b[k])]
    print(json.dumps({**lines, "arrays": len(a.files), "differing": differing, "bit_identical": not differing}))
    sys.exit(1 if differing else 0)


if __name__ == "__main__":
    main()
