"""CPU: host logic of the reference-API mirror, the Keras-HDF5 loader and the C-ABI surface."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import GOLDEN, REFERENCE, ROOT, pkg


# ---- C ABI -----------------------------------------------------------------------------------
def test_abi_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "vae21.h")).read()
    declared = set(re.findall(r"\b(vae21_[a-z0-9_]+)\s*\(", hdr))
    assert {"vae21_predict", "vae21_chi2", "vae21_set_model", "vae21_set_norm", "vae21_forward_normalised"} <= declared
    L = pkg("_lib")
    lib = L.load()
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/vae21.h but not exported"
    assert declared == set(L.EXPORTS)
    assert lib.vae21_version() == 100


def test_no_cpu_fallback_without_gpu():
    L = pkg("_lib")
    if L.device_count() > 0:
        pytest.skip("GPU present")
    with pytest.raises(L.Vae21Error):
        L.Handle(0)
    emu = pkg("emulator")
    pp = pkg("preprocess")
    e = emu.DirectEmulator(stats=pp.NormStats(np.zeros(7), np.ones(7), np.zeros(451, np.float32), np.float32(1)))
    with pytest.raises(L.Vae21Error):  # fails loudly, never computes on the CPU
        e.predict(np.ones(7))


def test_the_product_never_touches_the_oracle():
    """oracle/ is test infrastructure: nothing the package ships (Python, CUDA/C++ sources, the import shim, the header) may
    import, include, load or even name it -- only tests/, bench.py's CPU legs and __graft_entry__ (build + smoke check) do."""
    shipped = []
    for top in ("21cmvae_b200", "VeryAccurateEmulator", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, top)):
            shipped += [os.path.join(dirpath, f) for f in files if f.endswith((".py", ".cu", ".cuh", ".h", ".c", ".cpp"))]
    assert len(shipped) > 15
    pat = re.compile(r"\boracle\b|refmath|train_ref|mcmc_ref|chain_fp32")
    for path in shipped:
        hits = [ln.strip() for ln in open(path, encoding="utf-8", errors="replace") if pat.search(ln)]
        # the only mentions allowed are comments/docstrings saying that a CPU oracle EXISTS for a kernel
        assert not [h for h in hits if re.search(r"\b(import|include|CDLL|dlopen|open)\b", h)], (path, hits)


def test_null_and_state_errors_are_reported_not_crashes():
    L = pkg("_lib")
    lib = L.load()
    assert lib.vae21_create(0, None) == 1  # VAE21_ERR_ARG
    assert b"null" in lib.vae21_last_error()
    assert lib.vae21_predict(None, None, 0, 0, 0, None, 0, 0, None) != 0
    assert lib.vae21_destroy(None) == 0


# ---- Keras HDF5 ------------------------------------------------------------------------------
def test_tiny_keras_file_loads():
    kh = pkg("keras_h5")
    w = kh.load_dense_chain(os.path.join(GOLDEN, "tiny_keras.h5"))
    exp = np.load(os.path.join(GOLDEN, "tiny_keras_expected.npz"))
    assert w.dims == [7, 16, 24, 11] and w.relu == [True, True, False]
    assert w.layer_names == ["em_hidden_layer_0", "em_hidden_layer_1", "dense_16"]
    for i in range(3):
        assert np.array_equal(w.kernels[i], exp[f"k{i}"]) and np.array_equal(w.biases[i], exp[f"b{i}"])


def test_save_load_roundtrip(tmp_path, rm):
    kh = pkg("keras_h5")
    ks, bs, relu = rm.glorot_chain(rm.DIRECT_DIMS, seed=4)
    p = str(tmp_path / "m.h5")
    kh.save_dense_chain(p, kh.DenseChainWeights(ks, bs, relu, name="emulator"))
    w = kh.load_dense_chain(p)
    assert w.dims == list(rm.DIRECT_DIMS) and w.n_params() == 371907  # notebooks/sample_notebook.ipynb:68
    assert all(np.array_equal(a, b) for a, b in zip(w.kernels, ks))
    assert all(np.array_equal(a, b) for a, b in zip(w.biases, bs))


def test_invalid_model_path_raises_ioerror(tmp_path):
    kh = pkg("keras_h5")
    with pytest.raises(IOError):
        kh.load_dense_chain(str(tmp_path / "missing.h5"))
    bad = tmp_path / "bad.h5"
    bad.write_bytes(b"not an hdf5 file at all")
    with pytest.raises(IOError):
        kh.load_dense_chain(str(bad))


def test_model_config_that_does_not_cover_the_weight_layers_is_refused(tmp_path):
    """A file whose model_config names other layers than model_weights must not load with guessed activations."""
    import json

    kh = pkg("keras_h5")
    h5 = pkg("h5lite")
    cfg = {"class_name": "Sequential", "config": {"name": "m", "layers": [
        {"class_name": "Dense", "config": {"name": "other_0", "units": 3, "activation": "relu"}},
        {"class_name": "Dense", "config": {"name": "other_1", "units": 2, "activation": "linear"}}]}}
    wr = h5.Writer()
    wr.set_attr("/", "keras_version", "2.7.0")
    wr.set_attr("/", "model_config", json.dumps(cfg))
    wr.create_group("/model_weights")
    wr.set_attr("/model_weights", "layer_names", ["dense_0", "dense_1"])
    for n, (i, o) in zip(("dense_0", "dense_1"), ((4, 3), (3, 2))):
        wr.create_dataset(f"/model_weights/{n}/{n}/kernel:0", np.zeros((i, o), np.float32))
        wr.create_dataset(f"/model_weights/{n}/{n}/bias:0", np.zeros((o,), np.float32))
        wr.set_attr(f"/model_weights/{n}", "weight_names", [f"{n}/kernel:0", f"{n}/bias:0"])
    p = str(tmp_path / "inconsistent.h5")
    wr.save(p)
    with pytest.raises(IOError):
        kh.load_dense_chain(p)


@pytest.mark.skipif(not os.path.isdir(REFERENCE), reason="reference checkout not present")
def test_reads_the_shipped_reference_models(ae_golden):
    kh = pkg("keras_h5")
    base = REFERENCE + "/VeryAccurateEmulator/models/autoencoder_based_emulator/"
    em = kh.load_dense_chain(base + "ae_emulator.h5")
    de = kh.load_dense_chain(base + "decoder.h5")
    assert em.dims == [7, 352, 352, 352, 224, 9] and de.dims == [9, 32, 352, 451]
    assert em.n_params() == 332425  # notebooks/sample_notebook.ipynb:423
    ch = em.concat(de)
    assert all(np.array_equal(a, b) for a, b in zip(ch.kernels, ae_golden["kernels"]))
    enc = kh.load_dense_chain(base + "encoder.h5")
    assert enc.dims == [451, 352, 9]
    ae = kh.load_dense_chain(base + "autoencoder.h5")
    assert ae.dims == [451, 352, 9, 32, 352, 451] and ae.n_params() == 333420


# ---- reference-API mirror (tests/test_emulator.py, tests/test_preprocess.py of the reference) ----
def test_gen_model():
    emu = pkg("emulator")
    hidden = [32, 64, 256]
    model = emu._gen_model(7, hidden, 451, "relu")
    all_dims = hidden + [451]
    assert len(model.layers) == len(all_dims)
    for i, layer in enumerate(model.layers):
        assert layer.output_shape[-1] == all_dims[i]
    assert emu._gen_model(7, emu.hidden_dims, 451, "relu").count_params() == 371907


def test_z_nu():
    emu = pkg("emulator")
    assert np.isclose(30, emu.freq2redshift(emu.redshift2freq(30)))
    nu = emu.redshift2freq(emu.redshifts)
    assert np.isclose(nu[0], 1420.4057517667 / 6) and np.isclose(nu[-1], 1420.4057517667 / 51)
    keep = nu.copy()
    emu.freq2redshift(nu)
    assert np.array_equal(nu, keep)  # no in-place mutation


def test_error_metric(rm):
    emu = pkg("emulator")
    rng = np.random.default_rng(0)
    s = rng.normal(size=(12, 451)) * 50
    assert np.allclose(emu.error(s, s), 0)
    t = s + rng.normal(size=s.shape)
    nu = emu.redshift2freq(emu.redshifts)
    for kw in [{}, {"relative": False}, {"flow": 50, "fhigh": 100}, {"flow": 50}, {"fhigh": 100, "relative": False}]:
        a = emu.error(s, t, nu_arr=nu, **kw)
        b = rm.error(s, t, nu_arr=nu, **kw)
        assert a.shape == b.shape and np.allclose(a, b)
    with pytest.raises(ValueError):
        emu.error(s, t, flow=50)


def test_relative_mse_loss():
    emu = pkg("emulator")
    pp = pkg("preprocess")
    rng = np.random.default_rng(1)
    st = (rng.normal(size=(40, 451)) * 30).astype(np.float32)
    y_true = pp.preproc(st[:10], st)
    y_pred = pp.preproc(st[-10:], st)
    mse = np.mean((y_true - y_pred) ** 2, axis=1)
    amp = np.max(np.abs(st[:10] / np.std(st)), axis=1)
    assert np.allclose(emu.relative_mse_loss(st)(y_true, y_pred), mse / amp**2, rtol=1e-5)


def test_preprocess_mirror_matches_reference_outputs():
    pp = pkg("preprocess")
    d = np.load(os.path.join(GOLDEN, "preprocess.npz"))
    assert np.array_equal(pp.par_transform(d["params64"], d["par_train"]), d["pt64"])
    assert np.array_equal(pp.par_transform(d["params32"], d["par_train"]), d["pt32"])
    assert np.array_equal(pp.par_transform(d["params64"][0], d["par_train"]), d["pt_single"])
    assert np.array_equal(pp.preproc(d["sig"], d["sig_train"]), d["pre"])
    assert np.array_equal(pp.unpreproc(d["sig"], d["sig_train"]), d["unpre"])
    st = pp.NormStats.from_training_set(d["par_train"], d["sig_train"])
    assert np.array_equal(pp.par_transform_stats(d["params64"], st), d["pt64"])
    keep = d["params64"].copy()
    pp.par_transform(d["params64"], d["par_train"])
    assert np.array_equal(keep, d["params64"])  # inputs are never mutated
    # tests/test_preprocess.py:12-26 of the reference
    t = pp.par_transform(d["par_train"], d["par_train"])
    assert np.allclose(t.min(axis=0), -1) and np.allclose(t.max(axis=0), 1)
    pre = pp.preproc(d["sig_train"], d["sig_train"])
    assert np.allclose(pre.mean(axis=0), 0, atol=1e-3)
    assert np.allclose(pp.unpreproc(pre, d["sig_train"]), d["sig_train"], atol=5e-5)


def test_direct_emulator_construction_contract(rm):
    emu = pkg("emulator")
    with pytest.raises(IOError):
        emu.DirectEmulator()  # no dataset anywhere, nothing is downloaded
    par = rm.draw_params(50, 1)
    sig = np.random.default_rng(2).normal(size=(50, 451)).astype(np.float32)
    e = emu.DirectEmulator(par_train=par, signal_train=sig)
    assert e.par_labels == ["fstar", "Vc", "fx", "tau", "alpha", "nu_min", "Rmfp"]
    assert e.emulator.input_dim == 7 and e.emulator.output_dim == 451
    assert [l.units for l in e.emulator.layers] == [288, 352, 288, 224, 451]
    assert np.allclose(e.frequencies, emu.redshift2freq(e.redshifts))
    with pytest.raises(NotImplementedError):
        e.save()
    with pytest.raises(IOError):
        e.load_model("/nonexistent/emulator.h5")
    e2 = emu.DirectEmulator(par_train=par, signal_train=sig, redshifts=None, frequencies=np.array([50.0, 100.0]))
    assert np.allclose(e2.redshifts, emu.NU_0 / (np.array([50.0, 100.0]) * 1e6) - 1)
    lines = []
    e.emulator.summary(print_fn=lines.append)
    assert any("371,907" in ln for ln in lines)


def test_load_model_takes_architecture_from_file(rm):
    emu = pkg("emulator")
    pp = pkg("preprocess")
    e = emu.DirectEmulator(stats=pp.NormStats(np.zeros(7), np.ones(7), np.zeros(11, np.float32), np.float32(1)))
    e.load_model(os.path.join(GOLDEN, "tiny_keras.h5"))
    assert [l.units for l in e.emulator.layers] == [16, 24, 11]
    assert len(e.emulator.get_weights()) == 6


# ---- optimiser state of a saved model (training_config + optimizer_weights, as tf.keras.Model.save writes them) ----
def test_optimizer_state_round_trips_through_keras_h5(tmp_path, rm):
    kh = pkg("keras_h5")
    tr = pkg("training")
    ks, bs, relu = rm.glorot_chain([7, 16, 12, 5], seed=4)
    w = kh.DenseChainWeights(ks, bs, relu, ["hidden_0", "hidden_1", "out"], name="emulator")
    rng = np.random.default_rng(0)
    opt = tr.Adam(learning_rate=0.0123, beta_1=0.85, beta_2=0.97, epsilon=1e-6)
    opt.iterations = 4321
    opt.m = rng.normal(size=w.n_params()).astype(np.float32)
    opt.v = rng.uniform(size=w.n_params()).astype(np.float32)
    path = str(tmp_path / "with_opt.h5")
    kh.save_dense_chain(path, w, optimizer=opt)
    w2 = kh.load_dense_chain(path)
    assert all(np.array_equal(a, b) for a, b in zip(w.kernels + w.biases, w2.kernels + w2.biases))
    st = kh.load_optimizer_state(path, w2)
    assert st is not None and st.iterations == 4321 and st.loss == "loss_function"
    assert st.learning_rate == float(np.float32(0.0123)) and st.beta_1 == float(np.float32(0.85))
    assert st.beta_2 == float(np.float32(0.97)) and st.epsilon == 1e-6
    assert np.array_equal(st.m, opt.m) and np.array_equal(st.v, opt.v)
    back = tr.Adam.from_state(st)
    assert back.iterations == 4321 and np.array_equal(back.m, opt.m) and back.learning_rate == opt.learning_rate
    # the file keeps Keras' names and shapes: iter is a 0-d int64, slots sit under Adam/<layer>/{kernel,bias}/{m,v}:0
    f = pkg("h5lite").File(path)
    names = [n.decode() if isinstance(n, bytes) else str(n) for n in np.asarray(f["optimizer_weights"].attrs["weight_names"]).ravel()]
    assert names[0] == "Adam/iter:0" and names[1] == "Adam/hidden_0/kernel/m:0" and names[-1] == "Adam/out/bias/v:0"
    it = f["optimizer_weights"]["Adam"]["iter:0"].read()
    assert it.shape == () and it.dtype == np.int64
    assert f["optimizer_weights"]["Adam"]["hidden_1"]["kernel"]["v:0"].read().shape == (16, 12)
    # without an optimiser nothing is written and nothing is found
    kh.save_dense_chain(path, w)
    assert kh.load_optimizer_state(path) is None
    # moments of another model are refused
    opt.m = opt.m[:-1]
    with pytest.raises(ValueError):
        kh.save_dense_chain(path, w, optimizer=opt)


@pytest.mark.skipif(not os.path.isdir(REFERENCE), reason="reference checkout not present")
def test_reads_the_optimizer_state_of_the_shipped_models():
    """ae_emulator.h5 was saved by Keras 2.7 after training: Adam(lr 2.78e-4 after ReduceLROnPlateau), slots for its 5 layers."""
    kh = pkg("keras_h5")
    base = REFERENCE + "/VeryAccurateEmulator/models/autoencoder_based_emulator/"
    w = kh.load_dense_chain(base + "ae_emulator.h5")
    st = kh.load_optimizer_state(base + "ae_emulator.h5", w)
    assert st is not None and st.loss == "mean_squared_error" and st.iterations > 0
    assert abs(st.learning_rate - 0.00027812839834950864) < 1e-12 and st.epsilon == 1e-7
    assert st.m.shape == st.v.shape == (w.n_params(),) and np.all(st.v >= 0) and np.any(st.m != 0)
    assert kh.load_optimizer_state(base + "decoder.h5") is None  # saved without compile


def test_load_model_restores_the_compiled_optimizer(tmp_path, rm):
    """tf.keras.models.load_model returns the model compiled as saved (emulator.py:334-337); save_model writes that state."""
    emu = pkg("emulator")
    pp = pkg("preprocess")
    kh = pkg("keras_h5")
    tr = pkg("training")
    ks, bs, relu = rm.glorot_chain([7, 8, 451], seed=9)
    pmin, pmax = rm.prior_par_stats()
    stats = pp.NormStats(pmin, pmax, np.zeros(451, np.float32), np.float32(1.0))
    e = emu.DirectEmulator(stats=stats)
    e.emulator = emu.DenseModel(kh.DenseChainWeights(ks, bs, relu, name="emulator"))
    opt = tr.Adam(0.005)
    opt.iterations, opt.m, opt.v = 77, np.full(e.emulator.weights.n_params(), 0.5, np.float32), np.full(e.emulator.weights.n_params(), 0.25, np.float32)
    e.emulator.compile(optimizer=opt, loss=None)
    path = str(tmp_path / "m.h5")
    e.save_model(path)
    e2 = emu.DirectEmulator(stats=stats)
    e2.load_model(path)
    got = e2.emulator._compiled["optimizer"]
    assert isinstance(got, tr.Adam) and got.iterations == 77 and got.learning_rate == float(np.float32(0.005))
    assert np.array_equal(got.m, opt.m) and np.array_equal(got.v, opt.v)


class _OnlyDLPack:
    """A buffer that speaks nothing but the DLPack protocol (what a jax array or another framework's tensor looks like)."""

    def __init__(self, a):
        self._a = a

    def __dlpack__(self, stream=None):
        return self._a.__dlpack__()

    def __dlpack_device__(self):
        return self._a.__dlpack_device__()


def test_dlpack_buffers_are_unwrapped_without_a_copy():
    """north_star: 'numpy/DLPack buffers'.  The binding reads pointer / shape / dtype straight out of the DLPack capsule."""
    L = pkg("_lib")
    a = np.arange(21, dtype=np.float64).reshape(3, 7)
    ptr, on_dev, shape, dt, dev, keep = L._unwrap(_OnlyDLPack(a))
    assert (ptr, on_dev, shape, dt, dev) == (a.ctypes.data, False, (3, 7), np.dtype(np.float64), None) and keep is not None
    import torch

    t = torch.arange(12, dtype=torch.float32).reshape(3, 4)[1:]       # non-zero storage offset
    ptr, on_dev, shape, dt, _, _ = L._unwrap(_OnlyDLPack(t))
    assert (ptr, on_dev, shape, dt) == (t.data_ptr(), False, (2, 4), np.dtype(np.float32))
    with pytest.raises(ValueError, match="C-contiguous"):
        L._unwrap(_OnlyDLPack(a[:, ::2]))
    with pytest.raises(TypeError, match="float32/float64"):
        L._unwrap(_OnlyDLPack(np.arange(6).reshape(2, 3)))
    with pytest.raises(TypeError):
        L._unwrap(object())
    # the emulator front end hands a host DLPack producer to numpy without copying it
    emu = pkg("emulator")
    p = emu.DirectEmulator._as_param_array(_OnlyDLPack(a))
    assert isinstance(p, np.ndarray) and p.ctypes.data == a.ctypes.data and p.shape == (3, 7)


def test_the_h5py_branch_of_the_loader_on_an_h5py_shaped_file_object(monkeypatch, trained_fixture, tmp_path):
    """north_star: weights are read 'via h5py'.  h5py is absent from this image (the built-in reader serves instead), so the
    branch that calls it is driven here with a stand-in exposing exactly the h5py calls the loader makes -- `File(path, "r")`,
    `.attrs[...]`, `in`, `group[name]`, `dataset[()]`, `.close()` -- whose `read()` is removed so that a call meant for the
    built-in reader fails.  Same weights and the same optimiser state must come out of both branches."""
    kh = pkg("keras_h5")
    h5lite = pkg("h5lite")
    closed = []

    class _DS:
        def __init__(self, d):
            self._d, self.attrs, self.shape = d, d.attrs, d.shape

        def __getitem__(self, key):
            assert key == ()
            return self._d.read()

    class _Grp:
        def __init__(self, g):
            self._g, self.attrs = g, g.attrs

        def __contains__(self, k):
            return k in self._g

        def __getitem__(self, k):
            o = self._g[k]
            return _Grp(o) if isinstance(o, h5lite.Group) else _DS(o)

        def close(self):
            closed.append(1)

    class _FakeH5py:
        @staticmethod
        def File(path, mode="r"):
            assert mode == "r"
            return _Grp(h5lite.File(path))

    ref_w = kh.load_dense_chain(trained_fixture["path"])
    # a file with optimiser state, written by this repository's writer
    tr = pkg("training")
    opt = tr.Adam(0.003)
    opt.iterations = 17
    opt.m = np.linspace(-1, 1, ref_w.n_params()).astype(np.float32)
    opt.v = np.linspace(0, 2, ref_w.n_params()).astype(np.float32)
    path = str(tmp_path / "with_state.h5")
    kh.save_dense_chain(path, ref_w, optimizer=opt)
    ref_state = kh.load_optimizer_state(path)
    monkeypatch.setattr(kh, "h5py", _FakeH5py)
    monkeypatch.setattr(kh, "HAVE_H5PY", True)
    w = kh.load_dense_chain(trained_fixture["path"])
    assert w.layer_names == ref_w.layer_names and list(w.relu) == list(ref_w.relu) and w.name == ref_w.name
    for a, b in zip(w.kernels + w.biases, ref_w.kernels + ref_w.biases):
        assert np.array_equal(a, b)
    st = kh.load_optimizer_state(path)
    assert st.iterations == 17 == ref_state.iterations and st.learning_rate == ref_state.learning_rate
    assert np.array_equal(st.m, ref_state.m) and np.array_equal(st.v, ref_state.v) and np.array_equal(st.m, opt.m)
    assert len(closed) >= 3          # every h5py file the loader opened was closed again
    with pytest.raises(IOError):
        kh.load_dense_chain(str(tmp_path / "missing.h5"))


def test_the_binding_stub_of_integration_md_binds_the_real_library(rm):
    """INTEGRATION.md section B shows the file a maintainer of the reference would add.  Execute that very text against the built
    library: every symbol it names must exist with the argument counts of include/vae21.h, and without a GPU its first call must
    fail with the library's message, not crash."""
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    blocks = re.findall(r"```python\n(.*?)```", text, flags=re.S)
    stub = next(b for b in blocks if "VeryAccurateEmulator/_vae21.py" in b)
    L = pkg("_lib")
    stub = stub.replace('C.CDLL("libvae21.so")', f"C.CDLL({L.LIB_PATH!r})")
    ns = {}
    exec(compile(stub, "INTEGRATION.md:_vae21.py", "exec"), ns)
    hdr = open(os.path.join(ROOT, "include", "vae21.h")).read()
    for name in ("vae21_create", "vae21_set_model", "vae21_set_norm", "vae21_predict"):
        proto = re.search(rf"\b{name}\s*\((.*?)\);", hdr, flags=re.S).group(1)
        assert len(getattr(ns["_lib"], name).argtypes) == proto.count(",") + 1, name
    ks, bs, relu = rm.glorot_chain((7, 16, 451), seed=1)
    par = rm.draw_params(64, seed=2)
    sig = np.random.default_rng(3).normal(size=(64, 451)).astype(np.float32)
    if L.device_count() == 0:
        with pytest.raises(RuntimeError, match="CUDA"):
            ns["Handle"](ks, bs, [int(r) for r in relu], par, sig)
    else:
        h = ns["Handle"](ks, bs, [int(r) for r in relu], par, sig)
        got = h.predict(par[:5])
        want = rm.predict(par[:5], ks, bs, relu, *rm.par_stats(par), *rm.signal_stats(sig), squeeze=False)
        assert np.max(np.abs(got - want) / np.max(np.abs(want), axis=1, keepdims=True)) < 1e-5


def test_the_header_is_plain_c_and_links_against_the_library(tmp_path):
    """include/vae21.h is the drop-in boundary: it must compile as strict C99 (and C++11), and a C program using it must link
    against libvae21.so and get an error code -- not a crash -- from a call that cannot succeed."""
    import shutil
    import subprocess

    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no C compiler")
    L = pkg("_lib")
    src = tmp_path / "t.c"
    src.write_text('#include <stdio.h>\n#include "vae21.h"\n'
                   "int main(void) {\n"
                   "  vae21_handle* h = 0;\n"
                   "  if (vae21_version() != VAE21_VERSION) return 10;\n"
                   "  if (vae21_predict(0, 0, VAE21_F64, 0, 0, 0, 0, VAE21_FP32_SIMT, 0) == 0) return 11;\n"
                   "  if (vae21_last_error()[0] == 0) return 12;\n"
                   "  if (vae21_create(-1, &h) == 0 || h != 0) return 13;\n"
                   '  printf("%s\\n", vae21_last_error());\n'
                   "  return 0;\n}\n")
    inc = os.path.join(ROOT, "include")
    subprocess.run([gcc, "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-I", inc, str(src)], check=True)
    gxx = shutil.which("g++")
    if gxx:
        subprocess.run([gxx, "-std=c++11", "-fsyntax-only", "-x", "c++", "-I", inc, str(src)], check=True)
    exe = tmp_path / "t"
    libdir = os.path.dirname(L.LIB_PATH)
    subprocess.run([gcc, "-std=c99", "-I", inc, str(src), "-o", str(exe), "-L", libdir, "-l:" + os.path.basename(L.LIB_PATH),
                    "-Wl,-rpath," + libdir], check=True)
    p = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert p.returncode == 0, (p.returncode, p.stdout, p.stderr)
    assert p.stdout.strip()
