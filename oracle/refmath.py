"""CPU ORACLE -- test infrastructure only, never a product path.

Restatement of the reference's batched emulator evaluation
(``DirectEmulator.predict``) in plain numpy.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl
reference`` legs may import this module; the product (``21cmvae_b200``)
never does and fails loudly if its CUDA library is missing.

Parity status: PARTIALLY PINNED.
  * prologue/epilogue (``par_transform``/``unpreproc``): pinned against the
    reference's real ``preprocess.py`` (imported by file path in
    ``tests/golden/make_golden.py`` and ``tests/test_oracle.py``).
  * dense chain: the reference delegates it to TensorFlow/Keras
    (un-vendored, unpinned: requirements.txt:8 "tensorflow", README
    ">=2.5", shipped models written by Keras 2.7.0), which is absent from
    this image.  Pinned instead against float64 known-answer vectors
    computed from the reference's own shipped weight files
    (models/autoencoder_based_emulator/{ae_emulator,decoder}.h5, SURVEY.md
    section 8c) -- see ``tests/golden/``.  No TF ``predict`` output vector
    exists anywhere in the reference, so bit-level parity with TF itself is
    unpinned (DESIGN.md says so too).  What that freedom is worth is
    bounded: ``dense_chain_fp32_ordered`` evaluates the chain under eight
    explicit float32 summation orders (all within 2e-6 of the amplitude of
    the float64 arbiter, tests/test_oracle.py), its ``seq_fma`` order is
    reproduced bit for bit by the plain-C chain ``oracle/chain_fp32.c``, and
    the FP32 CUDA kernel reproduces that C chain bit for bit on a B200
    (tests/test_zz_fp32_vs_c_oracle.py).

Each function cites the reference lines it follows (paths relative to
/root/reference).
"""

from __future__ import annotations

from typing import Sequence

import numpy as np

FX_FLOOR = 10 ** (-6)  # preprocess.py:76


def transform_columns(parameters: np.ndarray) -> np.ndarray:
    """log10 of columns 0-2 with fx==0 -> 1e-6, rest copied; float64 out.

    VeryAccurateEmulator/preprocess.py:71-86 (and :89-97 for the training set).
    """
    p = np.asarray(parameters)
    if p.ndim == 1:
        p = p[None, :]
    cols12 = p[:, :2].copy()
    fx = p[:, 2].copy()
    fx[fx == 0] = FX_FLOOR
    out = np.empty(p.shape)  # float64, preprocess.py:81
    with np.errstate(divide="ignore", invalid="ignore"):
        out[:, :2] = np.log10(cols12)
        out[:, 2] = np.log10(fx)
    out[:, 3:] = p[:, 3:]
    return out


def par_stats(params_train: np.ndarray):
    """(min, max) per column of the transformed training parameters, float64.

    VeryAccurateEmulator/preprocess.py:89-101.
    """
    t = transform_columns(params_train)
    return np.min(t, axis=0), np.max(t, axis=0)


def par_transform(parameters: np.ndarray, params_train: np.ndarray) -> np.ndarray:
    """VeryAccurateEmulator/preprocess.py:49-110 (same op order: -=, /=, *=2, -=1)."""
    pmin, pmax = par_stats(params_train)
    return par_transform_cached(parameters, pmin, pmax)


def par_transform_cached(parameters, pmin, pmax) -> np.ndarray:
    x = transform_columns(parameters)
    x -= pmin
    x /= pmax - pmin
    x *= 2
    x -= 1
    return x


def signal_stats(signal_train: np.ndarray):
    """(mean[451], std scalar) in the dtype numpy gives the reference.

    VeryAccurateEmulator/preprocess.py:44-45: ``np.std`` over ALL elements
    (ddof=0), ``np.mean(axis=0)`` per bin; float32 in -> float32 out.
    """
    return np.mean(signal_train, axis=0), np.std(signal_train)


def unpreproc(signal: np.ndarray, signal_train: np.ndarray) -> np.ndarray:
    """VeryAccurateEmulator/preprocess.py:27-46 (mul, then add: two roundings)."""
    mu, sd = signal_stats(signal_train)
    return unpreproc_cached(signal, mu, sd)


def unpreproc_cached(signal, mu, sd) -> np.ndarray:
    out = signal * sd
    out += mu
    return out


def dense_chain(x: np.ndarray, kernels: Sequence[np.ndarray], biases: Sequence[np.ndarray],
                relu: Sequence[bool], dtype=np.float64) -> np.ndarray:
    """Keras ``Sequential`` of ``Dense`` layers: h = act(h @ kernel + bias).

    Kernel is ``[in, out]`` row-major; activation on flagged layers only.
    VeryAccurateEmulator/emulator.py:37-47 (``_gen_model``), called at :402.
    ``dtype=np.float64`` is the arbiter; ``np.float32`` mimics TF's
    arithmetic type (summation order inside SGEMM is unspecified there too).
    """
    h = np.asarray(x, dtype=dtype)
    if h.ndim == 1:
        h = h[None, :]
    for W, b, r in zip(kernels, biases, relu):
        h = h @ np.asarray(W, dtype=dtype) + np.asarray(b, dtype=dtype)
        if r:
            h = np.maximum(h, 0)
    return h


FP32_ORDERS = {
    # name: (k order, fused multiply-add, k block (0 = none), bias as the accumulator's start value)
    "seq_fma": ("seq", True, 0, False),        # k ascending, one rounding per term: the order of this repository's FP32 kernel
    "seq_mul_add": ("seq", False, 0, False),   # product and sum rounded separately (no FMA contraction)
    "rev_fma": ("rev", True, 0, False),
    "perm_fma": ("perm", True, 0, False),      # a random order of k, drawn per layer from `seed`
    "kc128_fma": ("seq", True, 128, False),    # a GEBP-style k panel: panels summed from zero, then added to the output
    "kc32_mul_add": ("seq", False, 32, False),
    "bias_first_fma": ("seq", True, 0, True),  # fused MatMul+BiasAdd: the bias seeds the accumulator
    "tree": ("tree", False, 0, False),         # pairwise summation of the separately rounded products
}


def dense_chain_fp32_ordered(x, kernels, biases, relu, order: str = "seq_fma", seed: int = 0) -> np.ndarray:
    """The Dense chain of emulator.py:37-47 / :402 in float32 with an EXPLICIT summation order.

    The reference hands ``h @ kernel + bias`` to TensorFlow's CPU MatMul (Eigen or oneDNN SGEMM), whose accumulation order, k
    blocking and use of FMA are not specified and differ between builds and machines; TensorFlow is absent here.  This function
    spans the orders such a kernel can take (``FP32_ORDERS``), so that tests can bound how far ANY float32 evaluation of the chain
    -- TensorFlow's included -- lies from the float64 arbiter.  An fp32 FMA is emulated as float32(float64(acc) + a*w): the product
    of two float32 is exact in float64, the sum is rounded to 53 and then to 24 bits (differs from a true FMA only in rare
    double-rounding cases, by one ulp).
    """
    kind, fma, kc, bias_first = FP32_ORDERS[order]
    rng = np.random.default_rng(seed)
    h = np.asarray(x, dtype=np.float32)
    if h.ndim == 1:
        h = h[None, :]

    def accumulate(h32, W32, ids, acc):
        if fma:
            h64, W64 = h32.astype(np.float64), W32.astype(np.float64)
            for k in ids:
                acc = (acc.astype(np.float64) + h64[:, k:k + 1] * W64[k:k + 1, :]).astype(np.float32)
        else:
            for k in ids:
                acc = acc + h32[:, k:k + 1] * W32[k:k + 1, :]
        return acc

    for W, b, r in zip(kernels, biases, relu):
        W = np.asarray(W, np.float32)
        b = np.asarray(b, np.float32)
        K, N = W.shape
        ids = np.arange(K)
        if kind == "rev":
            ids = ids[::-1]
        elif kind == "perm":
            ids = rng.permutation(K)
        zero = np.zeros((h.shape[0], N), np.float32)
        if kind == "tree":
            parts = [h[:, k:k + 1] * W[k:k + 1, :] for k in ids]
            while len(parts) > 1:
                nxt = [parts[i] + parts[i + 1] for i in range(0, len(parts) - 1, 2)]
                if len(parts) % 2:
                    nxt.append(parts[-1])
                parts = nxt
            out = parts[0] + b
        else:
            out = np.broadcast_to(b, zero.shape).astype(np.float32) if bias_first else zero
            if kc:
                for s in range(0, K, kc):
                    out = out + accumulate(h, W, ids[s:s + kc], zero)
            else:
                out = accumulate(h, W, ids, out)
            if not bias_first:
                out = out + b
        h = np.maximum(out, np.float32(0)) if r else out
        assert h.dtype == np.float32
    return h


def predict(params, kernels, biases, relu, pmin, pmax, mu, sd, dtype=np.float64, squeeze=True):
    """DirectEmulator.predict restated: emulator.py:383-407.

    par_transform (fp64) -> cast to fp32 as Keras does -> chain in ``dtype``
    -> unpreproc (fp32 constants) -> squeeze rule (:404-407).
    With ``dtype=float64`` the chain and the epilogue run in float64 on the
    float32-rounded inputs: that is the arbiter both GPU paths are held to.
    """
    x64 = par_transform_cached(params, pmin, pmax)
    x32 = x64.astype(np.float32)  # Keras casts its input to float32
    y = dense_chain(x32, kernels, biases, relu, dtype=dtype)
    if dtype == np.float32:
        out = unpreproc_cached(y, np.asarray(mu, np.float32), np.float32(sd))
    else:
        out = unpreproc_cached(y, np.asarray(mu, np.float64), np.float64(np.float32(sd)))
    if squeeze and out.shape[0] == 1:
        return out[0, :]
    return out


def chi2(pred: np.ndarray, obs: np.ndarray, inv_sigma: np.ndarray) -> np.ndarray:
    """Sum_i ((pred_i - obs_i) * inv_sigma_i)^2 per row, float64.

    Not in the reference (callers do it in numpy); oracle for the fused
    likelihood epilogue.
    """
    r = (np.asarray(pred, np.float64) - np.asarray(obs, np.float64)) * np.asarray(inv_sigma, np.float64)
    return np.sum(r * r, axis=-1)


def error(true_signal, pred_signal, relative=True, nu_arr=None, flow=None, fhigh=None):
    """VeryAccurateEmulator/emulator.py:129-192 restated (incl. the (N,1) quirk)."""
    if (flow or fhigh) and nu_arr is None:
        raise ValueError("No frequency array is given, cannot compute error in specified frequency band.")
    pred_signal = np.asarray(pred_signal)
    true_signal = np.asarray(true_signal)
    if pred_signal.ndim == 1:
        pred_signal = pred_signal[None, :]
        true_signal = true_signal[None, :]
    f = None
    if flow and fhigh:
        f = np.argwhere((nu_arr >= flow) & (nu_arr <= fhigh))[:, 0]
    elif flow:
        f = np.argwhere(nu_arr >= flow)
    elif fhigh:
        f = np.argwhere(nu_arr <= fhigh)
    if f is not None:
        pred_signal = pred_signal[:, f]
        true_signal = true_signal[:, f]
    err = np.sqrt(np.mean((pred_signal - true_signal) ** 2, axis=1))
    if relative:
        err /= np.max(np.abs(true_signal), axis=1)
        err *= 100
    return err


# --------------------------------------------------------------------------
# Synthetic inputs shared by tests and bench (the prior
# ranges are literature values, flagged as an assumption because the dataset
# is absent from the reference checkout).
# --------------------------------------------------------------------------

PRIOR_LO = np.array([1e-4, 4.2, 1e-6, 0.04, 1.0, 0.1, 10.0])
PRIOR_HI = np.array([0.5, 100.0, 1e3, 0.2, 1.5, 3.0, 50.0])
LOG_COLS = (0, 1, 2)


def draw_params(n: int, seed: int, zero_fx_frac: float = 0.01, dtype=np.float64) -> np.ndarray:
    """(n,7) parameter vectors from the prior; a fraction of fx set to exactly 0."""
    rng = np.random.default_rng(seed)
    u = rng.random((n, 7))
    p = np.empty((n, 7))
    for j in range(7):
        if j in LOG_COLS:
            lo, hi = np.log10(PRIOR_LO[j]), np.log10(PRIOR_HI[j])
            p[:, j] = 10.0 ** (lo + u[:, j] * (hi - lo))
        else:
            p[:, j] = PRIOR_LO[j] + u[:, j] * (PRIOR_HI[j] - PRIOR_LO[j])
    if zero_fx_frac > 0:
        p[rng.random(n) < zero_fx_frac, 2] = 0.0
    return p.astype(dtype)


def prior_par_stats():
    """pmin/pmax of the transformed prior box (stand-in for par_train's span)."""
    lo = PRIOR_LO.copy()
    hi = PRIOR_HI.copy()
    for j in LOG_COLS:
        lo[j] = np.log10(lo[j])
        hi[j] = np.log10(hi[j])
    return lo, hi


def glorot_chain(dims: Sequence[int], seed: int):
    """Glorot-uniform kernels / small random biases for a Dense chain (what
    ``_gen_model`` builds, emulator.py:41-47; biases made non-zero so the
    bias path is exercised)."""
    rng = np.random.default_rng(seed)
    kernels, biases = [], []
    for i in range(len(dims) - 1):
        fan_in, fan_out = dims[i], dims[i + 1]
        lim = np.sqrt(6.0 / (fan_in + fan_out))
        kernels.append(rng.uniform(-lim, lim, size=(fan_in, fan_out)).astype(np.float32))
        biases.append(rng.uniform(-0.1, 0.1, size=(fan_out,)).astype(np.float32))
    relu = [True] * (len(dims) - 2) + [False]
    return kernels, biases, relu


DIRECT_DIMS = (7, 288, 352, 288, 224, 451)  # emulator.py:196 + 7 params + 451 bins


def synthetic_signal_stats(kernels, biases, relu, sd_mk: float = 50.0):
    """Stand-in (mu[451], sd) when the dataset is absent:
    mean = emulator output at mid-prior (x = 0) in sigma-units * sd, sd = 50 mK."""
    y0 = dense_chain(np.zeros((1, kernels[0].shape[0])), kernels, biases, relu)[0]
    return (y0 * sd_mk).astype(np.float32), np.float32(sd_mk)
