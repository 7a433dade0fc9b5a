"""Load / save Dense-chain weights in the Keras-2.x HDF5 layout.

Replaces ``tf.keras.models.load_model`` at
/root/reference/VeryAccurateEmulator/emulator.py:333-337 (and :690-699 for the
autoencoder-based emulator).  The architecture comes from the file's
``model_config`` attribute (not from ``hidden_dims``), exactly like the
reference; weights come from ``model_weights/<layer>/<weight_names>``.

The optimiser state Keras stores next to the weights (``training_config`` attribute + ``optimizer_weights`` group: Adam's
iteration count and slot variables) is read by ``load_optimizer_state`` and written by ``save_dense_chain(..., optimizer=)``,
so that a model retrained here continues where it stopped, as ``tf.keras.models.load_model`` + ``fit`` does in the reference.

``h5py`` is used when importable (the north-star loader); otherwise the
dependency-free reader in ``h5lite`` parses the same files.
"""

from __future__ import annotations

import json
import os
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

from . import h5lite

try:  # pragma: no cover - h5py is absent from the build image
    import h5py  # type: ignore

    HAVE_H5PY = True
except Exception:  # noqa: BLE001
    h5py = None
    HAVE_H5PY = False

_ACTIVATIONS = {"relu": 1, "linear": 0, None: 0}


@dataclass
class DenseChainWeights:
    """A pure Dense stack: ``h = act(h @ kernels[i] + biases[i])``."""

    kernels: List[np.ndarray]
    biases: List[np.ndarray]
    relu: List[bool]
    layer_names: List[str] = field(default_factory=list)
    name: str = "model"
    keras_version: str = ""

    @property
    def dims(self) -> List[int]:
        return [int(self.kernels[0].shape[0])] + [int(k.shape[1]) for k in self.kernels]

    def validate(self):
        if not self.kernels:
            raise IOError("model has no Dense layers")
        for i, (k, b) in enumerate(zip(self.kernels, self.biases)):
            if k.ndim != 2 or b.ndim != 1 or k.shape[1] != b.shape[0]:
                raise IOError(f"layer {i}: kernel {k.shape} / bias {b.shape} mismatch")
            if i and self.kernels[i - 1].shape[1] != k.shape[0]:
                raise IOError(f"layer {i}: input width {k.shape[0]} != previous output {self.kernels[i-1].shape[1]}")

    def n_params(self) -> int:
        return int(sum(k.size + b.size for k, b in zip(self.kernels, self.biases)))

    def concat(self, other: "DenseChainWeights", name: Optional[str] = None) -> "DenseChainWeights":
        """Chain two models (AE emulator -> decoder, emulator.py:789-790)."""
        out = DenseChainWeights(
            self.kernels + other.kernels,
            self.biases + other.biases,
            self.relu + other.relu,
            self.layer_names + other.layer_names,
            name or (self.name + "+" + other.name),
            self.keras_version,
        )
        out.validate()
        return out


@dataclass
class AdamState:
    """tf.keras.optimizers.Adam as a saved model carries it: hyper-parameters (``training_config``), ``Adam/iter:0`` and the
    slot variables ``Adam/<layer>/{kernel,bias}/{m,v}:0``, here flat in ``get_weights()`` order (per layer kernel, then bias)."""

    learning_rate: float = 0.001
    beta_1: float = 0.9
    beta_2: float = 0.999
    epsilon: float = 1e-7
    iterations: int = 0
    m: Optional[np.ndarray] = None
    v: Optional[np.ndarray] = None
    loss: Optional[str] = None


def _as_str(x) -> str:
    if isinstance(x, bytes):
        return x.decode("utf-8")
    return str(x)


def _open(path: str):
    if not os.path.isfile(path):
        raise IOError(f"No file or directory found at {path}")
    if HAVE_H5PY:
        return h5py.File(path, "r"), True
    return h5lite.File(path), False


def _read_ds(obj, via_h5py: bool) -> np.ndarray:
    return np.asarray(obj[()] if via_h5py else obj.read())


def load_dense_chain(path: str) -> DenseChainWeights:
    """Parse a Keras ``.h5`` full-model or weights file into a Dense chain.

    Raises ``IOError`` when the path is not a valid model (the reference's
    documented contract, emulator.py:331).
    """
    f, via = _open(path)
    try:
        attrs = f.attrs
        acts = {}
        model_name = "model"
        if "model_config" in attrs:
            cfg = json.loads(_as_str(attrs["model_config"]))
            model_name = cfg.get("config", {}).get("name", model_name)
            for layer in cfg.get("config", {}).get("layers", []):
                cls = layer.get("class_name")
                lcfg = layer.get("config", {})
                if cls == "InputLayer":
                    continue
                if cls != "Dense":
                    raise IOError(f"{path}: layer class {cls!r} unsupported (pure Dense chains only)")
                act = lcfg.get("activation", "linear")
                if act not in _ACTIVATIONS:
                    raise IOError(f"{path}: activation {act!r} unsupported (relu/linear only)")
                if not lcfg.get("use_bias", True):
                    raise IOError(f"{path}: Dense without bias unsupported")
                acts[lcfg.get("name")] = bool(_ACTIVATIONS[act])
        mw = f["model_weights"] if "model_weights" in f else f
        layer_names = [_as_str(n) for n in np.asarray(mw.attrs["layer_names"]).ravel()]
        kernels, biases, relu, names = [], [], [], []
        for ln in layer_names:
            g = mw[ln]
            wn = [_as_str(n) for n in np.asarray(g.attrs.get("weight_names", [])).ravel()] if "weight_names" in g.attrs else []
            if not wn:
                continue  # InputLayer
            k = b = None
            for w in wn:
                arr = _read_ds(g[w], via)
                if w.endswith("kernel:0") or w.endswith("kernel"):
                    k = arr
                elif w.endswith("bias:0") or w.endswith("bias"):
                    b = arr
            if k is None or b is None:
                raise IOError(f"{path}: layer {ln} lacks kernel/bias")
            kernels.append(np.ascontiguousarray(k, dtype=np.float32))
            biases.append(np.ascontiguousarray(b, dtype=np.float32))
            names.append(ln)
            relu.append(acts.get(ln, None))
        # A weights-only file (no model_config) carries no activations: assume relu on all but the last layer, the architecture of
        # every model the reference ships.  A file WITH a model_config whose layer names do not cover the weight layers is
        # inconsistent -- refuse it instead of silently replacing every activation.
        if any(r is None for r in relu):
            if acts:
                missing = [n for n, r in zip(names, relu) if r is None]
                raise IOError(f"{path}: model_config has no entry for weight layer(s) {missing}")
            relu = [True] * (len(kernels) - 1) + [False]
        out = DenseChainWeights(kernels, biases, [bool(r) for r in relu], names, model_name,
                                _as_str(attrs["keras_version"]) if "keras_version" in attrs else "")
        out.validate()
        return out
    except (KeyError, ValueError, h5lite.H5LiteError) as e:
        raise IOError(f"{path} is not a valid Keras Dense-chain model: {e}") from e
    finally:
        if via:
            f.close()


def load_optimizer_state(path: str, w: Optional[DenseChainWeights] = None) -> Optional[AdamState]:
    """The Adam state of a Keras full-model file, or None when the file has no ``training_config`` or was compiled with another
    optimiser (such a model still predicts; it cannot be retrained here without an explicit ``compile``).  ``w``: the file's
    weights (``load_dense_chain(path)``), read again when omitted."""
    if w is None:
        w = load_dense_chain(path)
    f, via = _open(path)
    try:
        if "training_config" not in f.attrs:
            return None
        tc = json.loads(_as_str(f.attrs["training_config"]))
        oc = tc.get("optimizer_config") or {}
        if oc.get("class_name") != "Adam":
            return None
        c = oc.get("config", {})
        if c.get("amsgrad"):
            return None
        loss = tc.get("loss")
        st = AdamState(float(c.get("learning_rate", 0.001)), float(c.get("beta_1", 0.9)), float(c.get("beta_2", 0.999)),
                       float(c.get("epsilon", 1e-7)), 0, None, None, loss if isinstance(loss, str) else None)
        if "optimizer_weights" not in f:
            return st
        g = f["optimizer_weights"]
        names = [_as_str(n) for n in np.asarray(g.attrs["weight_names"]).ravel()] if "weight_names" in g.attrs else []
        slots = {}
        for n in names:
            obj = g
            for part in n.split("/"):
                obj = obj[part]
            arr = _read_ds(obj, via)
            parts = n.split("/")
            if parts[-1].startswith("iter"):
                st.iterations = int(np.asarray(arr).ravel()[0])
            elif len(parts) >= 4 and parts[-1][:1] in ("m", "v"):
                slots[(parts[-3], parts[-2], parts[-1][:1])] = np.asarray(arr, np.float32)
        if slots:
            flat = {"m": [], "v": []}
            for ln, k, b in zip(w.layer_names, w.kernels, w.biases):
                for which in ("m", "v"):
                    sk, sb = slots.get((ln, "kernel", which)), slots.get((ln, "bias", which))
                    if sk is None or sb is None or sk.shape != k.shape or sb.shape != b.shape:
                        raise IOError(f"{path}: optimizer_weights lack Adam/{ln}/{{kernel,bias}}/{which}:0 of the model's shape")
                    flat[which] += [sk.ravel(), sb.ravel()]
            st.m, st.v = np.concatenate(flat["m"]), np.concatenate(flat["v"])
        return st
    except (KeyError, ValueError, h5lite.H5LiteError) as e:
        raise IOError(f"{path}: unreadable optimizer state: {e}") from e
    finally:
        if via:
            f.close()


def save_dense_chain(path: str, w: DenseChainWeights, optimizer=None, loss: str = "loss_function"):
    """Write the Keras-2.x layout (model_config + model_weights) with h5lite.

    Files written here load back with ``load_dense_chain`` and follow the
    structure Keras 2.7 produces, so a TensorFlow
    installation can read them with ``load_model(..., compile=False)``.

    ``optimizer``: an object with Adam's attributes (``AdamState`` / ``training.Adam``); its hyper-parameters go to the
    ``training_config`` attribute and its iteration count and moments to ``optimizer_weights`` under the names Keras uses
    (``loss`` is the name recorded there: the reference's models carry "loss_function", emulator.py:51-83).
    """
    w.validate()
    dims = w.dims
    names = w.layer_names or [f"dense_{i}" for i in range(len(w.kernels))]
    in_name = names[0] + "_input"
    layers = [{
        "class_name": "InputLayer",
        "config": {"batch_input_shape": [None, dims[0]], "dtype": "float32", "sparse": False, "ragged": False,
                   "name": in_name},
        "name": in_name, "inbound_nodes": [],
    }]
    prev = in_name
    for n, k, r in zip(names, w.kernels, w.relu):
        layers.append({
            "class_name": "Dense",
            "config": {"name": n, "trainable": True, "dtype": "float32", "units": int(k.shape[1]),
                       "activation": "relu" if r else "linear", "use_bias": True,
                       "kernel_initializer": {"class_name": "GlorotUniform", "config": {"seed": None}},
                       "bias_initializer": {"class_name": "Zeros", "config": {}},
                       "kernel_regularizer": None, "bias_regularizer": None, "activity_regularizer": None,
                       "kernel_constraint": None, "bias_constraint": None},
            "name": n, "inbound_nodes": [[[prev, 0, 0, {}]]],
        })
        prev = n
    cfg = {"class_name": "Functional",
           "config": {"name": w.name, "layers": layers, "input_layers": [[in_name, 0, 0]],
                      "output_layers": [[prev, 0, 0]]}}
    wr = h5lite.Writer()
    wr.set_attr("/", "keras_version", w.keras_version or "2.7.0")
    wr.set_attr("/", "backend", "tensorflow")
    wr.set_attr("/", "model_config", json.dumps(cfg))
    wr.create_group("/model_weights")
    wr.set_attr("/model_weights", "layer_names", [in_name] + list(names))
    wr.set_attr("/model_weights", "backend", "tensorflow")
    wr.set_attr("/model_weights", "keras_version", w.keras_version or "2.7.0")
    wr.create_group(f"/model_weights/{in_name}")
    wr.set_attr(f"/model_weights/{in_name}", "weight_names", np.zeros((0,), dtype=np.float32))
    for n, k, b in zip(names, w.kernels, w.biases):
        wr.create_dataset(f"/model_weights/{n}/{n}/kernel:0", np.asarray(k, np.float32))
        wr.create_dataset(f"/model_weights/{n}/{n}/bias:0", np.asarray(b, np.float32))
        wr.set_attr(f"/model_weights/{n}", "weight_names", [f"{n}/kernel:0", f"{n}/bias:0"])
    if optimizer is not None:
        f32 = lambda x: float(np.float32(x))  # noqa: E731 (Keras stores float32 hyper-parameters)
        wr.set_attr("/", "training_config", json.dumps({
            "loss": loss, "metrics": None, "weighted_metrics": None, "loss_weights": None,
            "optimizer_config": {"class_name": "Adam", "config": {
                "name": "Adam", "learning_rate": f32(optimizer.learning_rate), "decay": 0.0, "beta_1": f32(optimizer.beta_1),
                "beta_2": f32(optimizer.beta_2), "epsilon": float(optimizer.epsilon), "amsgrad": False}}}))
        m, v = getattr(optimizer, "m", None), getattr(optimizer, "v", None)
        if m is not None and v is not None:
            if np.size(m) != w.n_params() or np.size(v) != w.n_params():
                raise ValueError("optimizer moments do not match the model's parameter count")
            wnames = ["Adam/iter:0"]
            wr.create_dataset("/optimizer_weights/Adam/iter:0", np.asarray(int(optimizer.iterations), dtype=np.int64))
            for which, flat in (("m", np.asarray(m, np.float32).ravel()), ("v", np.asarray(v, np.float32).ravel())):
                off = 0
                for n, k, b in zip(names, w.kernels, w.biases):
                    for part, ref in (("kernel", k), ("bias", b)):
                        wr.create_dataset(f"/optimizer_weights/Adam/{n}/{part}/{which}:0", flat[off:off + ref.size].reshape(ref.shape))
                        wnames.append(f"Adam/{n}/{part}/{which}:0")
                        off += ref.size
            wr.set_attr("/optimizer_weights", "weight_names", wnames)
    wr.save(path)
