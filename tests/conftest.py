import importlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
REFERENCE = "/root/reference"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def pkg(name=""):
    return importlib.import_module("21cmvae_b200" + ("." + name if name else ""))


@pytest.fixture(scope="session")
def rm():
    from oracle import refmath

    return refmath


@pytest.fixture(scope="session")
def ae_golden():
    d = np.load(os.path.join(GOLDEN, "ae_chain.npz"))
    n = sum(1 for k in d.files if k.startswith("k") and k[1:].isdigit())
    return {"x": d["x"], "y64": d["y64"], "latent64": d["latent64"], "relu": [bool(r) for r in d["relu"]],
            "kernels": [d[f"k{i}"] for i in range(n)], "biases": [d[f"b{i}"] for i in range(n)]}


@pytest.fixture(scope="session")
def direct_fixture(rm):
    """DirectEmulator architecture (7->288->352->288->224->451) with seeded Glorot weights and the
    synthetic normalisation constants of SURVEY.md section 8d (real weights/dataset are absent)."""
    ks, bs, relu = rm.glorot_chain(rm.DIRECT_DIMS, seed=2022)
    mu, sd = rm.synthetic_signal_stats(ks, bs, relu)
    pmin, pmax = rm.prior_par_stats()
    return {"kernels": ks, "biases": bs, "relu": relu, "mu": mu, "sd": sd, "pmin": pmin, "pmax": pmax}


@pytest.fixture(scope="session")
def trained_fixture():
    """DirectEmulator architecture with TRAINED-SCALE weights (|W| up to 2.0): the student of the reference's shipped autoencoder-based
    emulator, trained by this repository's CUDA trainer (tools/make_trained_fixture.py), plus the normalisation constants of its
    training set and 256 held-out (parameters, teacher signal) pairs."""
    kh = pkg("keras_h5")
    w = kh.load_dense_chain(os.path.join(GOLDEN, "direct_trained.h5"))
    d = np.load(os.path.join(GOLDEN, "direct_trained.npz"))
    return {"kernels": w.kernels, "biases": w.biases, "relu": [bool(r) for r in w.relu], "mu": d["sig_mean"], "sd": np.float32(d["sig_std"]),
            "pmin": d["par_min"], "pmax": d["par_max"], "par_test": d["par_test"], "signal_test": d["signal_test"],
            "path": os.path.join(GOLDEN, "direct_trained.h5")}


def have_gpu():
    try:
        return pkg("_lib").device_count() > 0
    except Exception:  # noqa: BLE001
        return False


_C_CHAIN_SRC = os.path.join(ROOT, "oracle", "chain_fp32.c")
_C_CHAIN_LIB = os.path.join(ROOT, "oracle", "_build", "libchain_fp32.so")


def build_c_oracle():
    """Compile oracle/chain_fp32.c (plain C, gcc) into oracle/_build/ unless an up-to-date library is there; None without gcc."""
    import shutil
    import subprocess

    if os.path.isfile(_C_CHAIN_LIB) and os.path.getmtime(_C_CHAIN_LIB) >= os.path.getmtime(_C_CHAIN_SRC):
        return _C_CHAIN_LIB
    gcc = shutil.which("gcc") or shutil.which("cc")
    if gcc is None:
        return None
    os.makedirs(os.path.dirname(_C_CHAIN_LIB), exist_ok=True)
    subprocess.run([gcc, "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", _C_CHAIN_LIB, _C_CHAIN_SRC, "-lm"], check=True)
    return _C_CHAIN_LIB


@pytest.fixture(scope="session")
def c_chain():
    """The plain-C float32 chain (k ascending, true fmaf, bias added afterwards): `f(x, kernels, biases, relu) -> (n, out) float32`."""
    import ctypes as C

    path = build_c_oracle()
    if path is None:
        pytest.skip("no C compiler for oracle/chain_fp32.c")
    lib = C.CDLL(path)
    lib.oracle_chain_fp32_seq_fma.argtypes = [C.c_void_p, C.c_long, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.oracle_chain_fp32_seq_fma.restype = C.c_int

    def run(x, kernels, biases, relu):
        x = np.ascontiguousarray(x, np.float32)
        ks = [np.ascontiguousarray(k, np.float32) for k in kernels]
        bs = [np.ascontiguousarray(b, np.float32) for b in biases]
        n_l = len(ks)
        dims = (C.c_int * (n_l + 1))(*([ks[0].shape[0]] + [k.shape[1] for k in ks]))
        kp = (C.c_void_p * n_l)(*[k.ctypes.data for k in ks])
        bp = (C.c_void_p * n_l)(*[b.ctypes.data for b in bs])
        rl = (C.c_int * n_l)(*[int(bool(r)) for r in relu])
        out = np.empty((len(x), ks[-1].shape[1]), np.float32)
        assert lib.oracle_chain_fp32_seq_fma(x.ctypes.data, len(x), n_l, dims, kp, bp, rl, out.ctypes.data) == 0
        return out

    return run
