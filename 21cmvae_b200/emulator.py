"""Host-side mirror of ``VeryAccurateEmulator.emulator`` for the hot path.

Same names, argument meaning and error behaviour as the reference
(/root/reference/VeryAccurateEmulator/emulator.py):

  DirectEmulator            :207-442   (predict :383-407 is THE hot path)
  _gen_model                :12-48     (defines the Dense stack)
  redshift2freq / freq2redshift / NU_0   :86-126
  error                     :129-192
  hidden_dims / redshifts   :196-197

What changed underneath: ``predict`` is one call into the CUDA library
(parameter transform, all Dense layers, de-normalisation fused in one
kernel); the training-set statistics are computed once per instance instead
of on every call; nothing is downloaded or opened at import time; no
TensorFlow.  There is no CPU fallback -- without the built library and a
B200 ``predict`` raises.
"""

from __future__ import annotations

import os
from typing import List, Optional, Sequence

import numpy as np

from . import _lib
from . import keras_h5
from . import preprocess as pp

PATH = os.path.dirname(os.path.abspath(__file__)) + "/"

NU_0 = 1420405751.7667  # Hz, rest frequency of the 21-cm line

# default parameters (emulator.py:196-197)
hidden_dims = [288, 352, 288, 224]
redshifts = np.linspace(5, 50, 451)

DATASET_ENV = "VAE21_DATASET"  # path to dataset_21cmVAE.h5 (keys par_train, signal_train, ...)
MODEL_ENV = "VAE21_MODEL"  # path to the DirectEmulator weights (models/emulator.h5 upstream)
PRECISION_ENV = "VAE21_PRECISION"  # fp32 | bf16x3 | fp16x3


def redshift2freq(z):
    """Redshift -> frequency in MHz."""
    return NU_0 / (1 + z) / 1e6


def freq2redshift(nu):
    """Frequency in MHz -> redshift.  (Unlike the reference, an ndarray argument
    is not modified in place -- emulator.py:124 multiplies the caller's array.)"""
    return NU_0 / (np.asarray(nu) * 1e6) - 1 if isinstance(nu, np.ndarray) else NU_0 / (nu * 1e6) - 1


def error(true_signal, pred_signal, relative=True, nu_arr=None, flow=None, fhigh=None):
    """Per-signal rms error (Eq. 1 of Bye et al. 2022), optionally inside a
    frequency band and optionally in % of the band's signal amplitude.

    Keeps the reference's behaviour including the result shape ``(N, 1)``
    when only one of ``flow`` / ``fhigh`` is given.
    """
    if (flow or fhigh) and nu_arr is None:
        raise ValueError("No frequency array is given, cannot compute error in specified frequency band.")
    pred_signal = np.asarray(pred_signal)
    true_signal = np.asarray(true_signal)
    if pred_signal.ndim == 1:
        pred_signal = pred_signal[None, :]
        true_signal = true_signal[None, :]
    sel = None
    if flow and fhigh:
        sel = np.argwhere((nu_arr >= flow) & (nu_arr <= fhigh))[:, 0]
    elif flow:
        sel = np.argwhere(nu_arr >= flow)
    elif fhigh:
        sel = np.argwhere(nu_arr <= fhigh)
    if sel is not None:
        pred_signal = pred_signal[:, sel]
        true_signal = true_signal[:, sel]
    err = np.sqrt(np.mean((pred_signal - true_signal) ** 2, axis=1))
    if relative:
        err /= np.max(np.abs(true_signal), axis=1)
        err *= 100
    return err


def relative_mse_loss(signal_train):
    """numpy form of the training loss (emulator.py:51-83): per-signal MSE divided
    by the squared amplitude of the (un-centred) true signal, in sigma units."""
    mean = np.mean(signal_train, axis=0) / np.std(signal_train)

    def loss_function(y_true, y_pred):
        y_true = np.asarray(y_true)
        y_pred = np.asarray(y_pred)
        amp = np.max(np.abs(y_true + mean), axis=1)
        return np.mean((y_true - y_pred) ** 2, axis=1) / amp**2

    return loss_function


# --------------------------------------------------------------------------------------------
# The object behind ``emu.emulator`` (a tf.keras.Model in the reference)
# --------------------------------------------------------------------------------------------


class _LayerView:
    def __init__(self, name, kernel, bias, activation):
        self.name = name
        self.units = int(kernel.shape[1])
        self.output_shape = (None, self.units)
        self.activation = activation
        self._k, self._b = kernel, bias

    def get_weights(self):
        return [self._k.copy(), self._b.copy()]


class DenseModel:
    """A stack of Dense layers evaluated by the CUDA library.

    Stands in for the ``tf.keras.Sequential`` the reference builds at
    emulator.py:37-47 for the calls users make on ``emu.emulator``:
    ``predict``, ``summary``, ``get_weights``, ``set_weights``, ``layers``,
    ``compile`` (recorded, only used by ``train``).
    """

    def __init__(self, weights: keras_h5.DenseChainWeights, device: int = 0):
        weights.validate()
        self.weights = weights
        self.name = weights.name
        self.device = int(device)
        self._handle: Optional[_lib.Handle] = None
        self._compiled = None

    # -- library handle, created on first use so that building a model needs no GPU
    @property
    def handle(self) -> _lib.Handle:
        if self._handle is None:
            h = _lib.Handle(self.device)
            h.set_model(self.weights.kernels, self.weights.biases, self.weights.relu)
            self._handle = h
        return self._handle

    @property
    def layers(self) -> List[_LayerView]:
        names = self.weights.layer_names or [f"dense_{i}" for i in range(len(self.weights.kernels))]
        return [_LayerView(n, k, b, "relu" if r else "linear")
                for n, k, b, r in zip(names, self.weights.kernels, self.weights.biases, self.weights.relu)]

    @property
    def input_dim(self):
        return self.weights.dims[0]

    @property
    def output_dim(self):
        return self.weights.dims[-1]

    def count_params(self):
        return self.weights.n_params()

    def get_weights(self):
        out = []
        for k, b in zip(self.weights.kernels, self.weights.biases):
            out += [k.copy(), b.copy()]
        return out

    def set_weights(self, arrays: Sequence[np.ndarray]):
        if len(arrays) != 2 * len(self.weights.kernels):
            raise ValueError("expected [kernel, bias] per layer")
        ks = [np.asarray(a, np.float32) for a in arrays[0::2]]
        bs = [np.asarray(a, np.float32) for a in arrays[1::2]]
        new = keras_h5.DenseChainWeights(ks, bs, list(self.weights.relu), list(self.weights.layer_names),
                                         self.weights.name, self.weights.keras_version)
        new.validate()
        self.weights = new
        if self._handle is not None:
            self._handle.set_model(ks, bs, new.relu)

    def compile(self, optimizer=None, loss=None, **kw):
        self._compiled = {"optimizer": optimizer, "loss": loss, **kw}

    def summary(self, print_fn=print):
        lines = [f'Model: "{self.name}"', "_" * 65, f"{'Layer (type)':<29}{'Output Shape':<26}{'Param #':<10}", "=" * 65]
        for lv in self.layers:
            n = lv._k.size + lv._b.size
            lines.append(f"{(lv.name + ' (Dense)'):<29}{str(lv.output_shape):<26}{n:<10}")
        total = self.count_params()
        lines += ["=" * 65, f"Total params: {total:,}", f"Trainable params: {total:,}", "Non-trainable params: 0", "_" * 65]
        for ln in lines:
            print_fn(ln)

    def predict(self, x, batch_size=None, verbose=0, precision=None, out=None, **_):
        """Dense stack on already-normalised inputs -> sigma-unit outputs, float32 ``(N, out)``.
        ``batch_size`` is accepted for Keras compatibility and ignored (one fused launch)."""
        prec = _resolve_precision(precision)
        if _is_device_array(x):
            return self.handle.forward_normalised(x, out=out, precision=prec)
        x = np.ascontiguousarray(x, dtype=np.float32)
        if x.ndim == 1:
            x = x[None, :]
        return self.handle.forward_normalised(x, out=out, precision=prec)

    __call__ = predict


def _is_device_array(x) -> bool:
    return hasattr(x, "__cuda_array_interface__")


def _resolve_precision(p) -> int:
    if p is None:
        p = os.environ.get(PRECISION_ENV, "fp32")
    if isinstance(p, str):
        try:
            return _lib.PRECISIONS[p.lower()]
        except KeyError:
            raise ValueError(f"unknown precision {p!r}; choose from {sorted(_lib.PRECISIONS)}") from None
    return int(p)


def _gen_model(in_dim, hidden_dims, out_dim, activation_func, name=None, seed=None, device=0):
    """A new, randomly initialised Dense stack ``in_dim -> hidden_dims... -> out_dim``:
    Glorot-uniform kernels and zero biases (the Keras defaults the reference relies on),
    ``activation_func`` after every hidden layer, linear output."""
    if activation_func not in ("relu", "linear", None):
        raise ValueError(f"activation {activation_func!r} unsupported by the CUDA kernels (relu/linear)")
    if in_dim is None:
        raise ValueError("in_dim=None (a model that succeeds another model) is not supported; chain the weights instead")
    rng = np.random.default_rng(seed)
    dims = [int(in_dim)] + [int(d) for d in hidden_dims] + [int(out_dim)]
    ks, bs = [], []
    for a, b in zip(dims[:-1], dims[1:]):
        lim = np.sqrt(6.0 / (a + b))
        ks.append(rng.uniform(-lim, lim, size=(a, b)).astype(np.float32))
        bs.append(np.zeros(b, np.float32))
    relu = [activation_func == "relu"] * len(hidden_dims) + [False]
    names = ["dense"] + [f"dense_{i}" for i in range(1, len(ks))]
    return DenseModel(keras_h5.DenseChainWeights(ks, bs, relu, names, name or "sequential"), device=device)


def _load_dataset(path):
    """Read the six arrays of dataset_21cmVAE.h5 (emulator.py:198-204)."""
    keys = ["par_train", "par_val", "par_test", "signal_train", "signal_val", "signal_test"]
    if keras_h5.HAVE_H5PY:
        with keras_h5.h5py.File(path, "r") as hf:
            return {k: hf[k][:] for k in keys}
    from . import h5lite

    f = h5lite.File(path)
    return {k: f[k].read() for k in keys}


class DirectEmulator:
    """User interface of the direct emulator (7 astrophysical parameters -> 451-bin global signal).

    Constructor arguments are those of the reference (emulator.py:208-220).  The reference takes its
    defaults from the dataset it opens at import time; here the dataset is optional: pass the arrays,
    or set ``VAE21_DATASET`` to the HDF5 file, or pass ``stats=NormStats(...)`` when only the 466
    normalisation constants are known.  ``device`` selects the GPU.
    """

    def __init__(self, par_train=None, par_val=None, par_test=None, signal_train=None, signal_val=None,
                 signal_test=None, hidden_dims=hidden_dims, activation_func="relu", redshifts=redshifts,
                 frequencies=None, *, stats: Optional[pp.NormStats] = None, device: int = 0, precision=None):
        (self.par_train, self.par_val, self.par_test, self.signal_train, self.signal_val, self.signal_test,
         stats) = _resolve_training_set(par_train, par_val, par_test, signal_train, signal_val, signal_test, stats)
        self.par_labels = ["fstar", "Vc", "fx", "tau", "alpha", "nu_min", "Rmfp"]
        self.stats = stats
        self.device = int(device)
        self.precision = precision
        n_in = int(np.shape(stats.par_min)[-1])
        n_out = int(np.shape(stats.sig_mean)[-1])
        self.emulator = _gen_model(n_in, hidden_dims, n_out, activation_func, name="emulator", device=self.device)
        self._norm_for = None
        if frequencies is None:
            if redshifts is not None:
                frequencies = redshift2freq(redshifts)
        elif redshifts is None:
            redshifts = freq2redshift(frequencies)
        self.redshifts = redshifts
        self.frequencies = frequencies

    # -- weights -------------------------------------------------------------------------------
    def load_model(self, model_path=None):
        """Load a saved Keras model (``.h5``).  Default: ``$VAE21_MODEL`` or ``models/emulator.h5`` next
        to this file (the reference's default path).  Raises IOError for an invalid path/model."""
        if model_path is None:
            model_path = os.environ.get(MODEL_ENV) or (PATH + "models/emulator.h5")
        w = keras_h5.load_dense_chain(model_path)
        self.emulator = DenseModel(w, device=self.device)
        self._norm_for = None
        # tf.keras.models.load_model returns the model COMPILED as it was saved (optimiser, its slot variables, the loss passed
        # as custom object, emulator.py:334-337): restore that, so that load_model() + train() continues a training run
        st = keras_h5.load_optimizer_state(model_path, w)
        if st is not None:
            from . import training as tr

            self.emulator.compile(optimizer=tr.Adam.from_state(st), loss=None)

    def save_model(self, model_path):
        """Write the current weights in the Keras-2.x HDF5 layout (loadable by ``load_model``), with the optimiser state of a
        compiled model (``training_config`` + ``optimizer_weights``) like ``tf.keras.Model.save``."""
        compiled = self.emulator._compiled or {}
        opt = compiled.get("optimizer")
        keras_h5.save_dense_chain(model_path, self.emulator.weights, optimizer=opt if hasattr(opt, "beta_1") else None)

    def save(self):
        raise NotImplementedError("Not implemented yet.")

    def train(self, epochs, callbacks=[], verbose="tqdm", *, seed=None, distributed=False):  # noqa: B006 - reference signature
        """Train the emulator (emulator.py:339-381): batches of 256, Adam and the relative-MSE loss of
        ``relative_mse_loss(signal_train)``, validation on the validation set after every epoch.

        ``callbacks`` are objects of ``training.EarlyStopping`` / ``training.ReduceLROnPlateau`` (same constructor
        arguments as the tf.keras callbacks of notebooks/Training.ipynb; TensorFlow objects are not accepted -- there is no
        TensorFlow here).  The model must have been compiled, as in the reference's notebooks:
        ``self.emulator.compile(optimizer=training.Adam(0.01), loss=relative_mse_loss(signal_train))``; an uncompiled model,
        another loss or a non-Adam optimiser raises (Keras would train them; this trainer cannot).  ``verbose="tqdm"`` maps to one line per
        epoch.  Extra keywords: ``seed`` fixes the shuffling, ``distributed=True`` splits every batch over the ranks of an
        initialised torch.distributed group (gradient all-reduce over NCCL).

        Returns (loss, val_loss): per-epoch training and validation losses, like the reference."""
        from . import training as tr

        if self.par_train is None or self.signal_train is None:
            raise ValueError("training needs par_train / signal_train (this emulator was built from NormStats only)")
        x_train = pp.par_transform(self.par_train, self.par_train).astype(np.float32)
        y_train = pp.preproc(np.asarray(self.signal_train), self.signal_train).astype(np.float32)
        loss_fn_mean = (np.mean(self.signal_train, axis=0) / np.std(self.signal_train)).astype(np.float32)
        amp_w = lambda y: (1.0 / np.max(np.abs(y + loss_fn_mean), axis=1) ** 2).astype(np.float32)  # noqa: E731 (emulator.py:70-80)
        x_val = y_val = w_val = None
        if self.par_val is not None and self.signal_val is not None and len(self.par_val):
            x_val = pp.par_transform(self.par_val, self.par_train).astype(np.float32)
            y_val = pp.preproc(np.asarray(self.signal_val), self.signal_train).astype(np.float32)
            w_val = amp_w(y_val)
        # Keras refuses to fit a model that was never compiled, and trains whatever loss / optimiser compile() was given.
        # This trainer implements exactly one configuration -- Adam on the relative-MSE loss (notebooks/Training.ipynb cell 4) --
        # so anything else is an error here, never a silent substitution.
        compiled = self.emulator._compiled
        if compiled is None:
            raise RuntimeError("You must compile your model before training/testing. Use `emulator.compile(optimizer, loss)` "
                               "with optimizer=training.Adam(lr) and loss=relative_mse_loss(signal_train).")
        loss = compiled.get("loss")
        if loss is not None and getattr(loss, "__name__", None) != "loss_function":
            raise ValueError(f"unsupported loss {loss!r}: the CUDA trainer implements relative_mse_loss(signal_train) "
                             "(emulator.py:51-83) only; pass that (or loss=None for the same objective)")
        opt = compiled.get("optimizer")
        if not isinstance(opt, tr.Adam):
            raise TypeError(f"unsupported optimizer {opt!r}: pass training.Adam(learning_rate, beta_1, beta_2, epsilon)")
        w = self.emulator.weights
        flat, hist = tr.fit(w.dims, w.relu, tr.flatten_weights(w.kernels, w.biases), x_train, y_train, amp_w(y_train), x_val, y_val,
                            w_val, optimizer=opt, epochs=epochs, batch_size=256, callbacks=list(callbacks), seed=seed,
                            device=self.device, distributed=distributed, verbose=1 if verbose in ("tqdm", 1, 2) else 0)
        ks, bs = tr.unflatten_weights(flat, w.dims)
        arrays = []
        for k, b in zip(ks, bs):
            arrays += [k, b]
        self.emulator.set_weights(arrays)
        self._norm_for = None
        self.last_training = hist
        return hist["loss"], hist["val_loss"]

    # -- evaluation ----------------------------------------------------------------------------
    def _handle(self) -> _lib.Handle:
        h = self.emulator.handle
        if self._norm_for is not h:
            s = self.stats
            mask = [1 if j in pp.LOG_COLUMNS else 0 for j in range(len(s.par_min))]
            h.set_norm(s.par_min, s.par_max, mask, 2, pp.FX_FLOOR, s.sig_mean, s.sig_std)
            self._norm_for = h
        return h

    @staticmethod
    def _as_param_array(params):
        if _is_device_array(params):
            return params
        if not isinstance(params, np.ndarray) and hasattr(params, "__dlpack__") and hasattr(params, "__dlpack_device__"):
            if int(params.__dlpack_device__()[0]) in (2, 13):  # kDLCUDA / kDLCUDAManaged: handed to the library as it is
                return params
            if not hasattr(params, "__array__"):
                params = np.from_dlpack(params)
        p = np.asarray(params)
        if p.ndim == 1:
            p = p[None, :]
        if p.dtype not in (np.float32, np.float64):
            # integer input: the reference's fx==0 -> 1e-6 substitution truncates back to 0 in an
            # integer array (emulator quirk) and log10 gives -inf; promote to
            # float64 instead so the floor applies.
            p = p.astype(np.float64)
        return np.ascontiguousarray(p)

    def predict(self, params, precision=None, out=None):
        """Predict global signal(s) [mK] from astrophysical parameters (order: ``par_labels``).

        ``params``: one 7-vector (list or 1-D array) or an ``(N, 7)`` array (float32/float64, host
        numpy or a CUDA tensor).  Returns float32 ``(451,)`` for exactly one row, else ``(N, 451)``
        (emulator.py:404-407).  The input is never modified; the result is a fresh array.
        """
        h = self._handle()
        p = self._as_param_array(params)
        pred = h.predict(p, out=out, precision=_resolve_precision(precision if precision is not None else self.precision))
        if pred.shape[0] == 1:
            return pred[0, :]
        return pred

    def chi2(self, params, observed, sigma, precision=None, return_argmin=False):
        """Fused likelihood: ``sum_k ((predict(params)[k] - observed[k]) / sigma[k])**2`` per row without
        materialising the spectra.  With ``return_argmin`` also returns ``(best_chi2, best_row)``."""
        h = self._handle()
        p = self._as_param_array(params)
        inv = 1.0 / np.broadcast_to(np.asarray(sigma, dtype=np.float64), (self.emulator.output_dim,))
        c, bv, bi = h.chi2(p, observed, inv.astype(np.float32), want_best=return_argmin,
                           precision=_resolve_precision(precision if precision is not None else self.precision))
        if return_argmin:
            return c, bv, bi
        return c

    def chi2_grid(self, points_per_dim, observed, sigma, first=0, count=None, precision=None, out=None):
        """chi^2 of ``observed`` over a regular grid spanning the training range of every parameter (the [-1, 1] box of
        ``preprocess.par_transform``: log-spaced in fstar, Vc, fx, linear in the others), generated on the GPU -- a 1e8-point
        grid never exists in memory.  Returns (best chi^2, flat grid index, the physical parameters of that point).  Shard large
        grids with ``first`` / ``count``.  Not in the reference."""
        h = self._handle()
        nd = len(self.stats.par_min)
        npts = [int(v) for v in (points_per_dim if np.ndim(points_per_dim) else [points_per_dim] * nd)]
        nout = len(self.stats.sig_mean)
        isig = 1.0 / np.broadcast_to(np.asarray(sigma, dtype=np.float64), (nout,))
        bv, bi = h.chi2_grid(npts, np.asarray(observed, np.float32), isig.astype(np.float32), first=first, count=count, out=out,
                             precision=_resolve_precision(precision if precision is not None else self.precision))
        return bv, bi, (self.grid_point(npts, bi) if bi >= 0 else None)

    def grid_point(self, points_per_dim, index):
        """Physical parameters of point ``index`` of the grid of ``chi2_grid`` (inverse of par_transform on the grid node)."""
        nd = len(self.stats.par_min)
        npts = [int(v) for v in (points_per_dim if np.ndim(points_per_dim) else [points_per_dim] * nd)]
        idx = np.unravel_index(int(index), npts)
        frac = np.array([i / (n - 1) if n > 1 else 0.0 for i, n in zip(idx, npts)])
        lo, hi = np.asarray(self.stats.par_min, np.float64), np.asarray(self.stats.par_max, np.float64)
        t = lo + frac * (hi - lo)
        return np.array([10.0 ** v if j in pp.LOG_COLUMNS else v for j, v in enumerate(t)])

    def test_error(self, relative=True, flow=None, fhigh=None, precision=None):
        """Error of the emulator for each signal in the test set (emulator.py:409-439), fused on the GPU: predict, band
        selection, rms and amplitude in one kernel (``vae21_error``); the 451-bin predictions are never copied back.
        Same results as ``error(self.signal_test, self.predict(self.par_test), ...)`` to float32 rounding, including the
        reference's result shape ``(N, 1)`` when only one of ``flow`` / ``fhigh`` is given (emulator.py:179-182)."""
        if self.par_test is None or self.signal_test is None:
            raise ValueError("no test set was given")
        return self.error_of(self.par_test, self.signal_test, relative=relative, flow=flow, fhigh=fhigh, precision=precision)

    def error_of(self, params, true_signal, relative=True, flow=None, fhigh=None, precision=None):
        """``error(true_signal, self.predict(params), ...)`` without materialising the predictions on the host."""
        nu = None if self.frequencies is None else np.asarray(self.frequencies)
        if (flow or fhigh) and nu is None:
            raise ValueError("No frequency array is given, cannot compute error in specified frequency band.")
        mask = None
        if flow and fhigh:
            mask = (nu >= flow) & (nu <= fhigh)
        elif flow:
            mask = nu >= flow
        elif fhigh:
            mask = nu <= fhigh
        p = self._as_param_array(params)
        if not _is_device_array(p) and p.ndim == 1:
            p = p[None, :]
        truth = true_signal if _is_device_array(true_signal) else np.ascontiguousarray(np.atleast_2d(true_signal), dtype=np.float32)
        err = self._handle().error(p, truth, None if mask is None else mask.astype(np.float32), relative=relative,
                                   precision=_resolve_precision(precision if precision is not None else self.precision))
        err = err.astype(np.float32)
        if bool(flow) != bool(fhigh):  # one-sided band: np.argwhere without [:, 0] keeps a trailing axis in the reference
            err = err[:, None]
        return err


# default parameters of the autoencoder-based emulator (emulator.py:521-525)
latent_dim = 9
enc_hidden_dims = [352]
dec_hidden_dims = [32, 352]
em_hidden_dims = [352, 352, 352, 224]


def _resolve_training_set(par_train, par_val, par_test, signal_train, signal_val, signal_test, stats):
    """The reference takes its defaults from the dataset it opens at import (emulator.py:198-204); here the arrays, or
    ``stats``, or the file named by ``VAE21_DATASET`` (else ``dataset_21cmVAE.h5`` next to this file) supply them."""
    if par_train is None and signal_train is None and stats is None:
        ds_path = os.environ.get(DATASET_ENV) or (PATH + "dataset_21cmVAE.h5")
        if not os.path.isfile(ds_path):
            raise IOError(
                "no training set: pass par_train/signal_train (or stats=NormStats), or point "
                f"{DATASET_ENV} at dataset_21cmVAE.h5 (nothing is downloaded at import, unlike the reference)")
        d = _load_dataset(ds_path)
        par_train, par_val, par_test = d["par_train"], d["par_val"], d["par_test"]
        signal_train, signal_val, signal_test = d["signal_train"], d["signal_val"], d["signal_test"]
    if stats is None:
        if par_train is None or signal_train is None:
            raise ValueError("both par_train and signal_train are needed to compute the normalisation constants")
        stats = pp.NormStats.from_training_set(par_train, signal_train)
    return par_train, par_val, par_test, signal_train, signal_val, signal_test, stats


class AutoEncoder:
    """Encoder + decoder pair of the autoencoder-based emulator (emulator.py:445-518): two Dense stacks held as
    ``DenseModel`` objects.  Only what ``AutoEncoderEmulator.predict`` needs runs on the GPU (the decoder, chained behind the
    emulator); calling the autoencoder on signals is not part of the hot path and is not implemented."""

    def __init__(self, signal_train=None, enc_hidden_dims=[], dec_hidden_dims=[], latent_dim=9, activation_func="relu", *,  # noqa: B006
                 n_out=None, device=0):
        if n_out is None:
            if signal_train is None:
                raise ValueError("signal_train (or n_out) is needed for the signal width")
            n_out = int(np.shape(signal_train)[-1])
        self.encoder = _gen_model(n_out, enc_hidden_dims, latent_dim, activation_func, name="encoder", device=device)
        # the reference builds the decoder with in_dim=None (input width inferred on first call, emulator.py:495-501)
        self.decoder = _gen_model(latent_dim, dec_hidden_dims, n_out, activation_func, name="decoder", device=device)

    def call(self, signals):
        raise NotImplementedError(
            "reconstructing signals with the autoencoder (451-wide inputs) is outside the evaluated hot path "
            "(SURVEY.md section 2: AutoEncoder is out of scope); only AutoEncoderEmulator.predict runs on the GPU")

    __call__ = call


class AutoEncoderEmulator:
    """Autoencoder-based emulator (emulator.py:528-842).  Same constructor and method signatures as the reference.

    ``predict`` evaluates two Keras models back to back in the reference (``emulator`` 7->...->latent, then the autoencoder's
    ``decoder`` latent->...->451, emulator.py:789-790).  Both are Dense chains, so they are concatenated and evaluated by the
    same fused kernel in one launch.  ``train`` and ``test_error(use_autoencoder=True)`` (which run the 451-input encoder) are
    outside the hot path and raise NotImplementedError.
    """

    AE_PATH = PATH + "models/autoencoder_based_emulator/"

    def __init__(self, par_train=None, par_val=None, par_test=None, signal_train=None, signal_val=None,
                 signal_test=None, latent_dim=latent_dim, enc_hidden_dims=enc_hidden_dims, dec_hidden_dims=dec_hidden_dims,
                 em_hidden_dims=em_hidden_dims, activation_func="relu", redshifts=redshifts, frequencies=None, *,
                 stats: Optional[pp.NormStats] = None, device: int = 0, precision=None):
        (self.par_train, self.par_val, self.par_test, self.signal_train, self.signal_val, self.signal_test,
         stats) = _resolve_training_set(par_train, par_val, par_test, signal_train, signal_val, signal_test, stats)
        self.par_labels = ["fstar", "Vc", "fx", "tau", "alpha", "nu_min", "Rmfp"]
        self.stats = stats
        self.device = int(device)
        self.precision = precision
        n_in = int(np.shape(stats.par_min)[-1])
        n_out = int(np.shape(stats.sig_mean)[-1])
        # untrained models of the requested shapes, as the reference builds them (emulator.py:640-665)
        self.autoencoder = AutoEncoder(None, enc_hidden_dims, dec_hidden_dims, latent_dim, activation_func, n_out=n_out,
                                       device=self.device)
        self.emulator = _gen_model(n_in, em_hidden_dims, latent_dim, activation_func, name="ae_emualtor", device=self.device)
        self._chain: Optional[DirectEmulator] = None
        self._chain_for = None
        if frequencies is None:
            if redshifts is not None:
                frequencies = redshift2freq(redshifts)
        elif redshifts is None:
            redshifts = freq2redshift(frequencies)
        self.redshifts = redshifts
        self.frequencies = frequencies

    @property
    def decoder(self) -> DenseModel:
        return self.autoencoder.decoder

    def load_model(self, emulator_path=None, encoder_path=None, decoder_path=None):
        """Load saved models (emulator.py:667-699); defaults are the files shipped with the reference.
        Raises IOError if a path does not point to a valid model."""
        em = keras_h5.load_dense_chain(emulator_path or (self.AE_PATH + "ae_emulator.h5"))
        en = keras_h5.load_dense_chain(encoder_path or (self.AE_PATH + "encoder.h5"))
        de = keras_h5.load_dense_chain(decoder_path or (self.AE_PATH + "decoder.h5"))
        if em.dims[-1] != de.dims[0] or en.dims[-1] != de.dims[0] or en.dims[0] != de.dims[-1]:
            raise IOError(f"models do not chain: emulator {em.dims}, encoder {en.dims}, decoder {de.dims}")
        self.emulator = DenseModel(em, device=self.device)
        self.autoencoder.encoder = DenseModel(en, device=self.device)
        self.autoencoder.decoder = DenseModel(de, device=self.device)
        self._chain = None

    def _fused_chain(self) -> DirectEmulator:
        key = (self.emulator, self.autoencoder.decoder)
        if self._chain is None or self._chain_for != key:
            chain = DirectEmulator(stats=self.stats, hidden_dims=[1], device=self.device, precision=self.precision,
                                   redshifts=self.redshifts, frequencies=self.frequencies)
            chain.emulator = DenseModel(self.emulator.weights.concat(self.autoencoder.decoder.weights, name="ae_emulator+decoder"),
                                        device=self.device)
            self._chain, self._chain_for = chain, key
        self._chain.par_test, self._chain.signal_test = self.par_test, self.signal_test
        self._chain.frequencies = self.frequencies
        return self._chain

    def train(self, epochs, ae_callbacks=[], em_callbacks=[], verbose="tqdm"):  # noqa: B006 - reference signature
        raise NotImplementedError(
            "training the autoencoder-based emulator (emulator.py:701-768) is outside the evaluated hot path; "
            "DirectEmulator.train is the supported retraining entry point")

    def predict(self, params, precision=None, out=None):
        """emulator.py:770-795: global signal(s) [mK] from astrophysical parameters; ``(451,)`` for exactly one row."""
        return self._fused_chain().predict(params, precision=precision if precision is not None else self.precision, out=out)

    def test_error(self, use_autoencoder=False, relative=True, flow=None, fhigh=None, precision=None):
        """emulator.py:797-842.  The emulator's error is fused on the GPU like DirectEmulator.test_error; the autoencoder's own
        error (``use_autoencoder=True``) needs the 451-input encoder and is not implemented."""
        if use_autoencoder:
            raise NotImplementedError("test_error(use_autoencoder=True) runs the 451-input encoder, which is outside the hot path")
        if self.par_test is None or self.signal_test is None:
            raise ValueError("no test set was given")
        return self._fused_chain().error_of(self.par_test, self.signal_test, relative=relative, flow=flow, fhigh=fhigh,
                                            precision=precision if precision is not None else self.precision)
