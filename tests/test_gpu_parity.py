"""GPU: parity of the CUDA paths with the CPU oracle, through the C-ABI (ctypes) and the
reference-API mirror.  Tolerances (BASELINE.json north_star):
  FP32-SIMT path      <= 1e-5 of the per-signal amplitude vs the float64 arbiter
  tensor-core paths   <= 0.01 mK rms and <= 0.05 mK max vs the float64 arbiter
"""
import os

import numpy as np
import pytest

from conftest import GOLDEN, pkg

pytestmark = pytest.mark.gpu

FP32_TOL = 1e-5        # of per-signal amplitude
TC_RMS_TOL_MK = 0.01   # mK
TC_MAX_TOL_MK = 0.05   # mK


def _direct(direct_fixture, device=0):
    emu = pkg("emulator")
    pp = pkg("preprocess")
    kh = pkg("keras_h5")
    f = direct_fixture
    e = emu.DirectEmulator(stats=pp.NormStats(f["pmin"], f["pmax"], f["mu"], f["sd"]), device=device)
    e.emulator = emu.DenseModel(kh.DenseChainWeights(f["kernels"], f["biases"], f["relu"], name="emulator"), device=device)
    return e


def _oracle(rm, f, params, **kw):
    return rm.predict(params, f["kernels"], f["biases"], f["relu"], f["pmin"], f["pmax"], f["mu"], f["sd"], **kw)


def _rel_err(got, want):
    amp = np.max(np.abs(want), axis=-1, keepdims=True)
    return float(np.max(np.abs(got.astype(np.float64) - want) / amp))


@pytest.fixture(scope="module")
def emu_direct(direct_fixture):
    return _direct(direct_fixture)


def tc_or_skip(e):
    if not e.emulator.handle.info()["tc_supported"]:
        pytest.skip("tensor-core kernel not available for this stack")


# ---- config 1: 1,024 synthetic vectors, seed 1024 ----------------------------------
def test_fp32_config1_matches_oracle(rm, direct_fixture, emu_direct):
    params = rm.draw_params(1024, seed=1024)
    want = _oracle(rm, direct_fixture, params)
    got = emu_direct.predict(params, precision="fp32")
    assert got.shape == (1024, 451) and got.dtype == np.float32 and got.flags["C_CONTIGUOUS"]
    assert _rel_err(got, want) <= FP32_TOL
    # and within fp32 noise of a float32 numpy evaluation (what TF's arithmetic type gives)
    want32 = _oracle(rm, direct_fixture, params, dtype=np.float32)
    assert _rel_err(got, want32.astype(np.float64)) <= FP32_TOL


@pytest.mark.parametrize("prec", ["bf16x3", "fp16x3", "fp16e4m3"])
def test_tc_config1_within_mk_tolerance(rm, direct_fixture, emu_direct, prec):
    tc_or_skip(emu_direct)
    params = rm.draw_params(1024, seed=1024)
    want = _oracle(rm, direct_fixture, params)
    got = emu_direct.predict(params, precision=prec).astype(np.float64)
    d = got - want
    assert np.sqrt(np.mean(d * d, axis=1)).max() <= TC_RMS_TOL_MK
    assert np.abs(d).max() <= TC_MAX_TOL_MK


# ---- reference test_predict (tests/test_emulator.py:55-69) --------------------------------------
@pytest.mark.parametrize("prec", ["fp32", "bf16x3"])
def test_single_vs_batched_and_squeeze(rm, direct_fixture, emu_direct, prec):
    if prec != "fp32":
        tc_or_skip(emu_direct)
    params = rm.draw_params(10, seed=5)
    single = emu_direct.predict(params[0], precision=prec)
    assert single.shape == (451,)
    assert emu_direct.predict(list(params[0]), precision=prec).shape == (451,)      # 1-D list input
    assert emu_direct.predict(params[:1], precision=prec).shape == (451,)           # (1, 7) squeezes too
    batched = emu_direct.predict(params, precision=prec)
    assert batched.shape == (10, 451)
    assert np.allclose(batched[0], single, atol=5e-5)


# ---- ragged sizes, empty input, fx == 0, float32 input -----------------------------------------
@pytest.mark.parametrize("n", [0, 1, 2, 63, 64, 65, 127, 128, 129, 300, 4097])
def test_fp32_ragged_sizes(rm, direct_fixture, emu_direct, n):
    params = rm.draw_params(n, seed=100 + n, zero_fx_frac=0.2).reshape(n, 7)
    h = emu_direct._handle()
    got = h.predict(np.ascontiguousarray(params), precision=0)
    assert got.shape == (n, 451)
    if n:
        want = _oracle(rm, direct_fixture, params, squeeze=False)
        assert _rel_err(got, want) <= FP32_TOL


@pytest.mark.parametrize("prec", [1, 3])
@pytest.mark.parametrize("n", [1, 127, 129, 1000])
def test_tc_ragged_sizes(rm, direct_fixture, emu_direct, n, prec):
    tc_or_skip(emu_direct)
    params = rm.draw_params(n, seed=200 + n, zero_fx_frac=0.2)
    got = emu_direct._handle().predict(params, precision=prec).astype(np.float64)
    d = got - _oracle(rm, direct_fixture, params, squeeze=False)
    assert np.sqrt(np.mean(d * d, axis=1)).max() <= TC_RMS_TOL_MK and np.abs(d).max() <= TC_MAX_TOL_MK


def test_fx_zero_uses_floor_and_inputs_unmodified(rm, direct_fixture, emu_direct):
    p = rm.draw_params(64, seed=7, zero_fx_frac=0.0)
    p[::2, 2] = 0.0
    keep = p.copy()
    got = emu_direct.predict(p)
    assert np.array_equal(p, keep)
    q = p.copy()
    q[::2, 2] = 1e-6
    assert np.array_equal(got, emu_direct.predict(q))  # bit-identical to passing the floor explicitly
    assert np.all(np.isfinite(got))


def test_float32_params_follow_numpy_float32_semantics(rm, direct_fixture, emu_direct):
    p32 = rm.draw_params(257, seed=12, zero_fx_frac=0.05).astype(np.float32)
    got = emu_direct.predict(p32)
    want = _oracle(rm, direct_fixture, p32)  # oracle takes log10 in float32 like numpy does
    assert _rel_err(got, want) <= FP32_TOL


def test_nonpositive_parameters_propagate_nan_like_numpy(rm, direct_fixture, emu_direct):
    p = rm.draw_params(4, seed=1)
    p[1, 0] = -1.0  # log10 of a negative number: numpy gives nan, no validation in the reference
    got = emu_direct.predict(p)
    assert np.all(np.isnan(got[1])) and np.all(np.isfinite(got[[0, 2, 3]]))


# ---- real trained weights: AE emulator + decoder chain (8 layers, 9-wide bottleneck) ------------
def test_ae_chain_forward_matches_float64_kats(ae_golden):
    emu = pkg("emulator")
    kh = pkg("keras_h5")
    g = ae_golden
    m = emu.DenseModel(kh.DenseChainWeights(g["kernels"], g["biases"], g["relu"], name="ae_chain"))
    y = m.predict(g["x"], precision="fp32")
    assert y.shape == g["y64"].shape
    assert _rel_err(y, g["y64"]) <= FP32_TOL
    # hand-recorded KAT of SURVEY.md 8c
    assert abs(float(y[0].astype(np.float64).sum()) - 9.912912209957) < 2e-3 and int(y[0].argmin()) == 89


@pytest.mark.parametrize("prec", ["bf16x3", "fp16e4m3"])
def test_ae_chain_tc_if_supported(ae_golden, prec):
    emu = pkg("emulator")
    kh = pkg("keras_h5")
    g = ae_golden
    m = emu.DenseModel(kh.DenseChainWeights(g["kernels"], g["biases"], g["relu"], name="ae_chain"))
    if not m.handle.info()["tc_supported"]:
        pytest.skip("AE chain does not fit the tensor-core plan")
    y = m.predict(g["x"], precision=prec).astype(np.float64)
    # sigma units; 50 mK per sigma => 0.01 mK rms = 2e-4 sigma, 0.05 mK max = 1e-3 sigma
    d = y - g["y64"]
    assert np.sqrt(np.mean(d * d, axis=1)).max() <= 2e-4 and np.abs(d).max() <= 1e-3


def test_load_model_end_to_end_from_h5(rm):
    emu = pkg("emulator")
    pp = pkg("preprocess")
    kh = pkg("keras_h5")
    w = kh.load_dense_chain(os.path.join(GOLDEN, "tiny_keras.h5"))
    pmin, pmax = rm.prior_par_stats()
    mu = np.linspace(-100, 0, 11).astype(np.float32)
    e = emu.DirectEmulator(stats=pp.NormStats(pmin, pmax, mu, np.float32(30)))
    e.load_model(os.path.join(GOLDEN, "tiny_keras.h5"))
    p = rm.draw_params(77, seed=4)
    want = rm.predict(p, w.kernels, w.biases, w.relu, pmin, pmax, mu, np.float32(30))
    assert _rel_err(e.predict(p), want) <= FP32_TOL


# ---- emu.emulator.predict (bare stack on normalised inputs) -------------------------------------
def test_forward_normalised(rm, direct_fixture, emu_direct):
    f = direct_fixture
    x = np.random.default_rng(3).uniform(-1, 1, size=(200, 7)).astype(np.float32)
    y = emu_direct.emulator.predict(x)
    want = rm.dense_chain(x, f["kernels"], f["biases"], f["relu"])
    assert y.shape == (200, 451) and _rel_err(y, want) <= FP32_TOL


# ---- fused chi^2 and argmin ----------------------------------------------------------------------
@pytest.mark.parametrize("prec", ["fp32", "bf16x3", "fp16x3", "fp16e4m3"])
def test_fused_chi2_and_argmin(rm, direct_fixture, emu_direct, prec):
    if prec != "fp32":
        tc_or_skip(emu_direct)
    params = rm.draw_params(5000, seed=77)
    pred = _oracle(rm, direct_fixture, params)
    rng = np.random.default_rng(7)
    truth = pred[1234] + rng.normal(size=451) * 0.5
    sigma = np.full(451, 25.0)
    want = rm.chi2(pred, truth.astype(np.float32), (1 / sigma).astype(np.float32))
    c, bv, bi = emu_direct.chi2(params, truth, sigma, precision=prec, return_argmin=True)
    assert c.shape == (5000,) and c.dtype == np.float32
    assert np.allclose(c, want, rtol=2e-4, atol=1e-6)
    assert bi == int(np.argmin(want)) == 1234 and np.isclose(bv, want[1234], rtol=2e-4)
    # argmin only (no chi2 array written)
    h = emu_direct._handle()
    none, bv2, bi2 = h.chi2(params, truth.astype(np.float32), (1 / sigma).astype(np.float32), want_chi2=False,
                            precision=pkg("_lib").PRECISIONS[prec])
    # every path is bitwise reproducible: the tensor-core kernel issues its MMAs in one fixed order (csrc/tc_kernel.cuh, issue table)
    assert none is None and bi2 == 1234 and bv2 == bv


# ---- device-resident buffers (torch / __cuda_array_interface__), async on the caller's stream ----
def test_device_resident_predict(rm, direct_fixture, emu_direct):
    import torch

    params = rm.draw_params(3000, seed=9)
    host = emu_direct.predict(params)
    t = torch.from_numpy(params).cuda()
    out = emu_direct.predict(t)
    assert isinstance(out, torch.Tensor) and out.is_cuda and out.shape == (3000, 451)
    torch.cuda.synchronize()
    assert np.array_equal(out.cpu().numpy(), host)  # same kernel, same bits
    pre = torch.empty((3000, 451), dtype=torch.float32, device="cuda")
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        emu_direct.predict(t, out=pre)
    s.synchronize()
    assert np.array_equal(pre.cpu().numpy(), host)


# ---- full BASELINE size: 1M rows, size-independent properties ------------------------------------
def test_one_million_rows_properties(rm, direct_fixture, emu_direct):
    import torch

    n = 1_000_000
    params = rm.draw_params(n, seed=20220322)
    t = torch.from_numpy(params).cuda()
    out = torch.empty((n, 451), dtype=torch.float32, device="cuda")
    emu_direct.predict(t, out=out, precision="fp32")
    torch.cuda.synchronize()
    assert bool(torch.isfinite(out).all())
    # (1) batch-split invariance: any sub-batch evaluated alone gives the same bits (rows are independent)
    for lo, hi in [(0, 1000), (499_937, 500_321), (n - 77, n)]:
        sub = emu_direct.predict(np.ascontiguousarray(params[lo:hi]), precision="fp32")
        assert np.array_equal(sub, out[lo:hi].cpu().numpy())
    # (2) spot rows against the oracle
    idx = np.random.default_rng(0).choice(n, 512, replace=False)
    want = _oracle(rm, direct_fixture, params[idx])
    assert _rel_err(out[torch.from_numpy(idx).cuda()].cpu().numpy(), want) <= FP32_TOL
    # (3) host-buffer pipeline (chunked copies) == device-resident launch, checksum of checksums
    host = emu_direct.predict(params[:200_000], precision="fp32")
    assert np.array_equal(host, out[:200_000].cpu().numpy())
    # (4) tensor-core path on the same million rows stays inside the mK tolerance on sampled rows
    if emu_direct.emulator.handle.info()["tc_supported"]:
        out_tc = torch.empty_like(out)
        emu_direct.predict(t, out=out_tc, precision="bf16x3")
        torch.cuda.synchronize()
        d = out_tc[torch.from_numpy(idx).cuda()].cpu().numpy().astype(np.float64) - want
        assert np.sqrt(np.mean(d * d, axis=1)).max() <= TC_RMS_TOL_MK and np.abs(d).max() <= TC_MAX_TOL_MK
        dd = (out_tc - out).abs().max().item()
        assert dd <= TC_MAX_TOL_MK


def test_kernel_launch_counter_moves(emu_direct, rm):
    h = emu_direct._handle()
    before = h.info()["kernel_launches"]
    emu_direct.predict(rm.draw_params(10, 1))
    assert h.info()["kernel_launches"] == before + 1


def test_tc_paths_are_bitwise_reproducible(rm, direct_fixture):
    """Every tcgen05.mma of a CTA pair is issued by ONE thread in program order (tc_kernel.cuh), so the accumulation order -- and
    with it every output bit -- is fixed: identical hashes within a process and across processes, in the DEFAULT configuration."""
    import subprocess
    import sys

    code = (
        "import sys, importlib, hashlib, numpy as np; sys.path.insert(0, %r); sys.path.insert(0, %r);"
        "from oracle import refmath as rm;"
        "emu=importlib.import_module('21cmvae_b200.emulator'); pp=importlib.import_module('21cmvae_b200.preprocess');"
        "kh=importlib.import_module('21cmvae_b200.keras_h5');"
        "ks,bs,relu=rm.glorot_chain(rm.DIRECT_DIMS, seed=2022); mu,sd=rm.synthetic_signal_stats(ks,bs,relu);"
        "pmin,pmax=rm.prior_par_stats(); e=emu.DirectEmulator(stats=pp.NormStats(pmin,pmax,mu,sd));"
        "e.emulator=emu.DenseModel(kh.DenseChainWeights(ks,bs,relu));"
        "p=rm.draw_params(20000, seed=5);"
        "h=[hashlib.sha256(e.predict(p, precision=q).tobytes()).hexdigest() for q in ('bf16x3','fp16e4m3') for _ in range(3)];"
        "print(h[0][:32]+h[3][:32] if len(set(h[:3]))==1 and len(set(h[3:]))==1 else 'DIFF')"
    ) % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ)
    env.pop("VAE21_TC_DETERMINISTIC", None)
    outs = [subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300).stdout.strip()
            for _ in range(2)]
    assert outs[0] == outs[1] and outs[0] != "DIFF" and len(outs[0]) == 64, outs


def test_autoencoder_emulator_end_to_end_from_h5(tmp_path, rm, ae_golden):
    """AutoEncoderEmulator (emulator.py:770-795): two Keras files -> one fused chain, real trained weights."""
    emu = pkg("emulator")
    pp = pkg("preprocess")
    kh = pkg("keras_h5")
    g = ae_golden
    em = kh.DenseChainWeights(g["kernels"][:5], g["biases"][:5], g["relu"][:5], name="AE_Emulator")
    de = kh.DenseChainWeights(g["kernels"][5:], g["biases"][5:], g["relu"][5:], name="Decoder")
    rng = np.random.default_rng(5)
    en = kh.DenseChainWeights([rng.normal(0, 0.05, (451, 352)).astype(np.float32), rng.normal(0, 0.05, (352, 9)).astype(np.float32)],
                              [np.zeros(352, np.float32), np.zeros(9, np.float32)], [True, False], name="Encoder")
    kh.save_dense_chain(str(tmp_path / "ae_emulator.h5"), em)
    kh.save_dense_chain(str(tmp_path / "encoder.h5"), en)
    kh.save_dense_chain(str(tmp_path / "decoder.h5"), de)
    pmin, pmax = rm.prior_par_stats()
    mu = np.linspace(-120, 10, 451).astype(np.float32)
    sd = np.float32(47.5)
    ae = emu.AutoEncoderEmulator(stats=pp.NormStats(pmin, pmax, mu, sd))
    # the reference's positional order: emulator, encoder, decoder (emulator.py:667-672)
    ae.load_model(str(tmp_path / "ae_emulator.h5"), str(tmp_path / "encoder.h5"), str(tmp_path / "decoder.h5"))
    assert ae.autoencoder.encoder.weights.dims == [451, 352, 9] and ae.autoencoder.decoder.weights.dims == [9, 32, 352, 451]
    p = rm.draw_params(333, seed=21)
    want = rm.predict(p, g["kernels"], g["biases"], g["relu"], pmin, pmax, mu, sd)
    got = ae.predict(p)
    assert got.shape == (333, 451) and _rel_err(got, want) <= FP32_TOL
    assert ae.predict(p[0]).shape == (451,)
    # latent head alone (what the reference's self.emulator.predict returns)
    x = rm.par_transform_cached(p, pmin, pmax).astype(np.float32)
    lat = ae.emulator.predict(x)
    assert lat.shape == (333, 9)
    assert np.allclose(lat, rm.dense_chain(x, g["kernels"][:5], g["biases"][:5], g["relu"][:5]), atol=2e-5)


def test_time_predict_and_pinned_pool(rm, direct_fixture, emu_direct):
    """vae21_time_predict (device-timed back-to-back launches) and the pinned host pool."""
    import torch

    L = pkg("_lib")
    h = emu_direct._handle()
    p = torch.from_numpy(rm.draw_params(50_000, seed=2)).cuda()
    o = torch.empty((50_000, 451), dtype=torch.float32, device="cuda")
    ms = h.time_predict(p, o, precision=L.FP32_SIMT, iters=3)
    assert 0 < ms < 100
    want = _oracle(rm, direct_fixture, p[:64].cpu().numpy())
    assert _rel_err(o[:64].cpu().numpy(), want) <= FP32_TOL
    a = L.pinned_empty((1000, 451), np.float32)
    addr = a.ctypes.data
    a[:] = 1.0
    del a
    import gc

    gc.collect()
    b = L.pinned_empty((1000, 451), np.float32)  # same size class: the pool hands the block back
    assert b.ctypes.data == addr


# ---- device-generated grid (BASELINE config 3) ------------------------------------------------------
@pytest.mark.parametrize("prec", ["fp32", "fp16e4m3"])
def test_chi2_grid_matches_explicit_grid(rm, direct_fixture, emu_direct, prec):
    """chi2_grid generates the grid nodes inside the kernel; the same nodes passed as explicit parameter rows must give the
    same chi^2 (the explicit path takes log10 of 10**t, so agreement is to rounding, not bitwise), in the reference's C order."""
    if prec != "fp32":
        tc_or_skip(emu_direct)
    npts = [3, 2, 4, 3, 2, 3, 5]
    total = int(np.prod(npts))
    f = direct_fixture
    axes = [np.linspace(f["pmin"][j], f["pmax"][j], n) for j, n in enumerate(npts)]
    mesh = np.stack(np.meshgrid(*axes, indexing="ij"), axis=-1).reshape(total, 7)
    params = mesh.copy()
    params[:, :3] = 10.0 ** params[:, :3]
    truth = _oracle(rm, f, params[1234 % total])
    obs = (truth + np.random.default_rng(5).normal(size=451) * 3).astype(np.float32)
    want, _, _ = emu_direct.chi2(params, obs, 25.0, precision=prec, return_argmin=True)
    import torch

    out = torch.empty(total, dtype=torch.float32, device="cuda")
    bv, bi, bp = emu_direct.chi2_grid(npts, obs, 25.0, precision=prec, out=out)
    got = out.cpu().numpy()
    assert np.allclose(got, want, rtol=2e-3, atol=1e-4)
    assert bi == int(np.argmin(got)) and bv == pytest.approx(float(got[bi]))
    assert np.allclose(bp, params[bi], rtol=1e-12)
    # a shard of the grid: same values, global index
    lo, cnt = 301, 500
    out2 = torch.empty(cnt, dtype=torch.float32, device="cuda")
    bv2, bi2, _ = emu_direct.chi2_grid(npts, obs, 25.0, first=lo, count=cnt, precision=prec, out=out2)
    got2 = out2.cpu().numpy()
    if prec == "fp32":
        assert np.array_equal(got2, got[lo:lo + cnt])
    else:  # the shard's rows sit at other tile positions than in the full launch; the per-row arithmetic is the same, a tolerance is kept
        assert np.allclose(got2, got[lo:lo + cnt], rtol=1e-5, atol=0)
    assert bi2 == lo + int(np.argmin(got2))


# ---- fused figure of merit (emulator.py:129-192, :409-439) -------------------------------------------
@pytest.mark.parametrize("prec", ["fp32", "fp16e4m3"])
def test_fused_error_matches_numpy_error(rm, direct_fixture, prec):
    """DirectEmulator.test_error runs predict + band selection + rms + amplitude in one kernel; it must agree with the
    reference's numpy `error` applied to the oracle's predictions, for every band variant incl. the (N, 1) quirk."""
    emu = pkg("emulator")
    f = direct_fixture
    par = rm.draw_params(333, seed=41)
    pred = _oracle(rm, f, par)
    truth = (pred * (1 + 0.01 * np.random.default_rng(2).normal(size=pred.shape)) + 0.3).astype(np.float32)
    e = _direct(f)
    if prec != "fp32":
        tc_or_skip(e)
    e.par_test, e.signal_test = par, truth
    nu = e.frequencies
    for kw in ({}, {"relative": False}, {"flow": 50.0, "fhigh": 120.0}, {"flow": 80.0}, {"fhigh": 100.0, "relative": False}):
        want = emu.error(truth, pred.astype(np.float32), nu_arr=nu, **{"relative": True, **kw})
        got = e.test_error(precision=prec, **kw)
        assert got.shape == want.shape, kw
        assert np.allclose(got, want, rtol=2e-4, atol=1e-5), kw
    # single signal, host float64 truth
    one = e.error_of(par[0], truth[0].astype(np.float64), precision=prec)
    assert one.shape == (1,) and np.isclose(one[0], emu.error(truth[0], pred[0].astype(np.float32))[0], rtol=2e-4)
    with pytest.raises(Exception):
        e.error_of(par, truth, flow=1e6, precision=prec)  # empty band


# ---- planner generality: other Dense stacks through every tensor-core format --------------------------------
@pytest.mark.parametrize("dims", [(7, 16, 11), (3, 40, 24, 451), (16, 100, 200, 100, 30), (7, 480, 64), (5, 33, 47, 19, 130, 7)])
@pytest.mark.parametrize("prec", ["bf16x3", "fp16x3", "fp16e4m3"])
def test_tc_paths_on_other_architectures(rm, dims, prec):
    """The schedule planner is generic over Dense stacks (widths not multiples of 16, narrow outputs, up to 16 inputs); whatever it
    accepts must meet the tensor-core tolerance in sigma units (0.01 mK rms / 0.05 mK max at 50 mK per sigma)."""
    emu = pkg("emulator")
    kh = pkg("keras_h5")
    ks, bs, relu = rm.glorot_chain(dims, seed=sum(dims))
    rng = np.random.default_rng(1)
    bs = [rng.normal(scale=0.1, size=b.shape).astype(np.float32) for b in bs]
    m = emu.DenseModel(kh.DenseChainWeights(ks, bs, relu, name="generic"))
    if not m.handle.info()["tc_supported"]:
        pytest.skip("stack does not fit the tensor-core plan")
    x = rng.uniform(-1, 1, size=(777, dims[0])).astype(np.float32)
    want = rm.dense_chain(x, ks, bs, relu)
    y32 = m.predict(x, precision="fp32")
    assert _rel_err(y32, want) <= FP32_TOL
    d = m.predict(x, precision=prec).astype(np.float64) - want
    assert np.sqrt(np.mean(d * d, axis=1)).max() <= 2e-4 and np.abs(d).max() <= 1e-3


# ---- trained-scale weights: the precision claims of the tensor-core formats, pinned where they are tight -------------------------
def _trained(trained_fixture, device=0):
    emu = pkg("emulator")
    pp = pkg("preprocess")
    f = trained_fixture
    e = emu.DirectEmulator(stats=pp.NormStats(f["pmin"], f["pmax"], f["mu"], f["sd"]), device=device)
    e.load_model(f["path"])  # through the Keras-HDF5 loader, like the reference's load_model (emulator.py:319-337)
    return e


@pytest.mark.parametrize("prec", ["bf16x3", "fp16x3", "fp16e4m3"])
def test_tc_formats_on_the_reference_trained_chain_20k_inputs(rm, ae_golden, prec, capsys):
    """20,000 uniform inputs through the reference's REAL trained weights (ae_emulator.h5 -> decoder.h5, 8 Dense layers, |W| up to
    2.4) on every tensor-core format, against the float64 arbiter: 0.01 mK rms / 0.05 mK max in sigma units (sigma = 50 mK, SURVEY
    8d).  tools/precision_study.py predicts (float64 simulation) 2.5e-5 / 8.8e-7 / 5.9e-5 sigma rms for bf16x3 / fp16x3 / fp16e4m3."""
    emu = pkg("emulator")
    kh = pkg("keras_h5")
    g = ae_golden
    m = emu.DenseModel(kh.DenseChainWeights(g["kernels"], g["biases"], g["relu"], name="ae_chain"))
    if not m.handle.info()["tc_supported"]:
        pytest.skip("tensor-core kernel not available for this stack")
    x = np.random.default_rng(20211).uniform(-1, 1, size=(20_000, 7)).astype(np.float32)
    want = rm.dense_chain(x, g["kernels"], g["biases"], g["relu"], dtype=np.float64)
    m.handle.tc_saturation(reset=True)
    got = m.predict(x, precision=prec).astype(np.float64)
    d = (got - want) * 50.0  # mK at sigma = 50 mK
    rms, mx = float(np.sqrt(np.mean(d * d, axis=1)).max()), float(np.abs(d).max())
    with capsys.disabled():
        print(f"\n[{prec}] real AE chain, 20,000 inputs: rms max {rms:.2e} mK (budget {TC_RMS_TOL_MK}, margin x{TC_RMS_TOL_MK / rms:.1f}), "
              f"max {mx:.2e} mK (budget {TC_MAX_TOL_MK}, margin x{TC_MAX_TOL_MK / mx:.1f})")
    assert rms <= TC_RMS_TOL_MK and mx <= TC_MAX_TOL_MK
    assert m.handle.tc_saturation() == 0  # max |pre-activation| of this chain is ~18: far inside the e4m3 / fp16 operand ranges


@pytest.mark.parametrize("prec", ["fp16e4m3", "bf16x3"])
def test_trained_direct_emulator_1m_rows(rm, trained_fixture, prec, capsys):
    """BASELINE config 2 on TRAINED-scale DirectEmulator weights, the default tensor-core format included: 1M prior draws, device
    resident; 4,096 sampled rows against the float64 arbiter inside 0.01 mK rms / 0.05 mK max (the fixture's own sigma = 45.8 mK);
    the operand-range counter stays 0; two launches agree bit for bit."""
    import torch

    e = _trained(trained_fixture)
    tc_or_skip(e)
    n = 1_000_000
    params = rm.draw_params(n, seed=20220322)
    t = torch.from_numpy(params).cuda()
    out = torch.empty((n, 451), dtype=torch.float32, device="cuda")
    h = e._handle()
    h.tc_saturation(reset=True)
    e.predict(t, out=out, precision=prec)
    torch.cuda.synchronize()
    idx = np.random.default_rng(1).choice(n, 4096, replace=False)
    want = _oracle(rm, trained_fixture, params[idx])
    d = out[torch.from_numpy(idx).cuda()].cpu().numpy().astype(np.float64) - want
    rms, mx = float(np.sqrt(np.mean(d * d, axis=1)).max()), float(np.abs(d).max())
    with capsys.disabled():
        print(f"\n[{prec}] trained DirectEmulator fixture, 1M rows: rms max {rms:.2e} mK (margin x{TC_RMS_TOL_MK / rms:.1f}), "
              f"max {mx:.2e} mK (margin x{TC_MAX_TOL_MK / mx:.1f})")
    assert rms <= TC_RMS_TOL_MK and mx <= TC_MAX_TOL_MK
    assert h.tc_saturation() == 0
    out2 = torch.empty_like(out)
    e.predict(t, out=out2, precision=prec)
    torch.cuda.synchronize()
    assert torch.equal(out, out2)
    # the FP32 parity path on the same weights (sampled rows)
    got32 = e.predict(params[idx], precision="fp32")
    assert _rel_err(got32, want) <= FP32_TOL


def test_trained_fixture_reproduces_the_reference_accuracy_tests(trained_fixture):
    """tests/test_emulator.py:55-69 of the reference, on the trained fixture: a single prediction has shape (451,), its relative rms
    error against the true (teacher) signal is a few per cent at most, and row 0 of a batched call equals the single call (atol
    5e-5 mK)."""
    emu = pkg("emulator")
    e = _trained(trained_fixture)
    f = trained_fixture
    pred = e.predict(f["par_test"][0])
    assert pred.shape == (451,)
    assert emu.error(f["signal_test"][0], pred)[0] < 5.0
    batch = e.predict(f["par_test"][:10])
    assert batch.shape == (10, 451)
    assert np.allclose(batch[0], pred, atol=5e-5)
    e.par_test, e.signal_test = f["par_test"], f["signal_test"]
    err = e.test_error()
    assert err.shape == (256,) and 0.3 < float(err.mean()) < 1.5


def test_operand_range_counter_fires_outside_the_range(rm, direct_fixture):
    """vae21_get_tc_stats: hidden activations beyond the e4m3 range (448) are counted for fp16e4m3 and not for bf16x3."""
    emu = pkg("emulator")
    kh = pkg("keras_h5")
    f = direct_fixture
    ks = [k.copy() for k in f["kernels"]]
    ks[0] = ks[0] * 6000.0  # first hidden layer far beyond 448
    m = emu.DenseModel(kh.DenseChainWeights(ks, f["biases"], f["relu"], name="hot"))
    if not m.handle.info()["tc_supported"]:
        pytest.skip("tensor-core kernel not available for this stack")
    x = np.random.default_rng(0).uniform(-1, 1, size=(1000, 7)).astype(np.float32)
    m.handle.tc_saturation(reset=True)
    m.predict(x, precision="bf16x3")
    assert m.handle.tc_saturation() == 0
    m.predict(x, precision="fp16e4m3")
    assert m.handle.tc_saturation(reset=True) > 0
    assert m.handle.tc_saturation() == 0
