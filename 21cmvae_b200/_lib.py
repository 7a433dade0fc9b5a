"""ctypes binding of the vae21 C-ABI library (include/vae21.h).

The library is built in-tree by ``__graft_entry__.build()`` (nvcc, sm_100a)
as ``21cmvae_b200/libvae21.so``.  There is no CPU fallback: if the library
is missing or no GPU is present, every compute call raises.
"""

from __future__ import annotations

import ctypes as C
import os
import threading
import weakref
from typing import Optional, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("VAE21_LIB", os.path.join(_HERE, "libvae21.so"))

F32, F64 = 0, 1
FP32_SIMT, TC_BF16X3, TC_FP16X3, TC_FP16E4M3 = 0, 1, 2, 3
PRECISIONS = {"fp32": FP32_SIMT, "fp32_simt": FP32_SIMT, "tc": TC_BF16X3, "bf16x3": TC_BF16X3,
              "tc_bf16x3": TC_BF16X3, "fp16x3": TC_FP16X3, "tc_fp16x3": TC_FP16X3,
              "fp16e4m3": TC_FP16E4M3, "tc_fp16e4m3": TC_FP16E4M3}

EXPORTS = [
    "vae21_version", "vae21_last_error", "vae21_device_count", "vae21_create", "vae21_destroy",
    "vae21_set_model", "vae21_set_norm", "vae21_predict", "vae21_forward_normalised", "vae21_chi2", "vae21_chi2_grid", "vae21_error", "vae21_mcmc_run", "vae21_check_plan",
    "vae21_host_alloc", "vae21_host_free", "vae21_host_trim", "vae21_get_info", "vae21_get_tc_stats", "vae21_time_predict",
    "vae21_trainer_create", "vae21_trainer_destroy", "vae21_trainer_num_params", "vae21_trainer_set_params",
    "vae21_trainer_get_params", "vae21_trainer_set_moments", "vae21_trainer_get_moments", "vae21_trainer_forward_backward", "vae21_trainer_adam", "vae21_trainer_epoch", "vae21_trainer_dp_begin", "vae21_trainer_dp_forward_backward", "vae21_trainer_dp_adam",
    "vae21_trainer_launches",
]


class Vae21Error(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"vae21 error {code}: {msg}")
        self.code = code


_lib = None
_lock = threading.Lock()


def load() -> C.CDLL:
    """Load libvae21.so (once).  Raises if it has not been built."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.isfile(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(this package has no CPU fallback)")
        lib = C.CDLL(LIB_PATH)
        vp, i32, i64, f32p = C.c_void_p, C.c_int, C.c_int64, C.POINTER(C.c_float)
        lib.vae21_version.restype = i32
        lib.vae21_last_error.restype = C.c_char_p
        lib.vae21_device_count.argtypes = [C.POINTER(i32)]
        lib.vae21_create.argtypes = [i32, C.POINTER(vp)]
        lib.vae21_destroy.argtypes = [vp]
        lib.vae21_set_model.argtypes = [vp, i32, C.POINTER(i32), C.POINTER(vp), C.POINTER(vp), C.POINTER(i32)]
        lib.vae21_set_norm.argtypes = [vp, i32, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(i32), i32,
                                       C.c_double, i32, f32p, C.c_float]
        lib.vae21_predict.argtypes = [vp, vp, i32, i32, i64, vp, i32, i32, vp]
        lib.vae21_forward_normalised.argtypes = [vp, vp, i32, i64, vp, i32, i32, vp]
        lib.vae21_chi2.argtypes = [vp, vp, i32, i32, i64, f32p, f32p, vp, i32, C.POINTER(C.c_float),
                                   C.POINTER(i64), i32, vp]
        lib.vae21_chi2_grid.argtypes = [vp, i32, C.POINTER(i32), C.POINTER(C.c_double), C.POINTER(C.c_double), i64, i64, f32p, f32p, vp,
                                        C.POINTER(C.c_float), C.POINTER(i64), i32, vp]
        lib.vae21_error.argtypes = [vp, vp, i32, i32, i64, vp, i32, f32p, i32, vp, i32, i32, vp]
        lib.vae21_host_alloc.argtypes = [C.c_size_t]
        lib.vae21_host_alloc.restype = vp
        lib.vae21_host_free.argtypes = [vp]
        lib.vae21_host_free.restype = None
        lib.vae21_host_trim.restype = None
        lib.vae21_get_info.argtypes = [vp, C.POINTER(i64), C.POINTER(C.c_float), C.POINTER(i32)]
        lib.vae21_get_tc_stats.argtypes = [vp, C.POINTER(i64), i32]
        lib.vae21_mcmc_run.argtypes = [vp, vp, vp, i64, i32, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_float),
                                       C.POINTER(C.c_float), C.c_double, C.c_uint64, i64, i32, i32, i32, vp, C.POINTER(i64)]
        lib.vae21_check_plan.argtypes = [i32, C.POINTER(i32), C.c_char_p, i32]
        lib.vae21_time_predict.argtypes = [vp, vp, i32, i64, vp, i32, i32, C.POINTER(C.c_float)]
        lib.vae21_trainer_create.argtypes = [i32, i32, C.POINTER(i32), C.POINTER(i32), i32, C.POINTER(vp)]
        lib.vae21_trainer_destroy.argtypes = [vp]
        lib.vae21_trainer_num_params.argtypes = [vp, C.POINTER(i64)]
        lib.vae21_trainer_set_params.argtypes = [vp, f32p, i32]
        lib.vae21_trainer_get_params.argtypes = [vp, f32p]
        lib.vae21_trainer_set_moments.argtypes = [vp, f32p, f32p]
        lib.vae21_trainer_get_moments.argtypes = [vp, f32p, f32p]
        lib.vae21_trainer_forward_backward.argtypes = [vp, vp, vp, vp, vp, i64, i32, C.c_float, vp, vp, vp]
        lib.vae21_trainer_adam.argtypes = [vp, vp, C.c_float, C.c_float, C.c_float, C.c_float, vp]
        lib.vae21_trainer_epoch.argtypes = [vp, vp, vp, vp, vp, i64, i32, C.c_float, C.c_float, C.c_float, C.c_float, i64, vp, vp]
        lib.vae21_trainer_dp_begin.argtypes = [vp, vp, vp, vp, vp, i64, i32, i32, i32, C.c_float, C.c_float, C.c_float, C.c_float, i64, vp, vp, vp]
        lib.vae21_trainer_dp_forward_backward.argtypes = [vp, vp]
        lib.vae21_trainer_dp_adam.argtypes = [vp, vp]
        lib.vae21_trainer_launches.argtypes = [vp, C.POINTER(i64)]
        for name in EXPORTS:  # a stale libvae21.so must fail here, not at first use
            getattr(lib, name)
        _lib = lib
        return lib


def _check(rc: int):
    if rc != 0:
        raise Vae21Error(rc, load().vae21_last_error().decode("utf-8", "replace"))


def check_plan(dims):
    """Host-only self-check of the tensor-core schedule of a Dense stack: (code, reason) with code 0 = consistent schedule,
    1 = the stack does not fit the tensor-core kernel, 2 = inconsistent schedule (a planner bug)."""
    lib = load()
    d = (C.c_int32 * len(dims))(*[int(v) for v in dims])
    buf = C.create_string_buffer(256)
    rc = lib.vae21_check_plan(len(dims) - 1, d, buf, 256)
    return rc, buf.value.decode("utf-8", "replace")


def device_count() -> int:
    n = C.c_int(0)
    rc = load().vae21_device_count(C.byref(n))
    return n.value if rc == 0 else 0


# ---- pinned host arrays ----------------------------------------------------------------------


def pinned_empty(shape, dtype=np.float32) -> np.ndarray:
    """A fresh numpy array in pinned host memory from the library's caching pool.
    The block returns to the pool when the array (and every view of it) is garbage collected."""
    lib = load()
    dtype = np.dtype(dtype)
    shape = tuple(int(s) for s in (shape if isinstance(shape, (tuple, list)) else (shape,)))
    nbytes = int(np.prod(shape, dtype=np.int64)) * dtype.itemsize
    ptr = lib.vae21_host_alloc(max(nbytes, 1))
    if not ptr:
        raise MemoryError(f"pinned allocation of {nbytes} bytes failed: {lib.vae21_last_error().decode()}")
    buf = (C.c_char * max(nbytes, 1)).from_address(ptr)
    weakref.finalize(buf, lib.vae21_host_free, ptr)
    arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape, dtype=np.int64))).reshape(shape)
    return arr


# ---- buffer unwrapping -----------------------------------------------------------------------


def _unwrap(obj, want_write=False) -> Tuple[int, bool, Tuple[int, ...], np.dtype, Optional[int], object]:
    """(pointer, on_device, shape, dtype, device_index, keepalive) for numpy arrays, torch tensors and
    anything exposing __cuda_array_interface__.  Requires C-contiguity."""
    if isinstance(obj, np.ndarray):
        if not obj.flags["C_CONTIGUOUS"]:
            raise ValueError("array must be C-contiguous")
        if want_write and not obj.flags["WRITEABLE"]:
            raise ValueError("output array must be writeable")
        return obj.ctypes.data, False, obj.shape, obj.dtype, None, obj
    cai = getattr(obj, "__cuda_array_interface__", None)
    if cai is not None:
        if cai.get("strides") is not None:
            # accept only C-contiguous strides
            shape = tuple(cai["shape"])
            item = np.dtype(cai["typestr"]).itemsize
            exp = []
            acc = item
            for s in reversed(shape):
                exp.append(acc)
                acc *= s
            if tuple(reversed(exp)) != tuple(cai["strides"]):
                raise ValueError("device array must be C-contiguous")
        dev = getattr(getattr(obj, "device", None), "index", None)
        return int(cai["data"][0]), True, tuple(cai["shape"]), np.dtype(cai["typestr"]), dev, obj
    if hasattr(obj, "__dlpack__") and hasattr(obj, "data_ptr"):  # torch CPU tensor
        t = obj
        if not t.is_contiguous():
            raise ValueError("tensor must be contiguous")
        import torch  # noqa: WPS433

        np_dt = {torch.float32: np.float32, torch.float64: np.float64}.get(t.dtype)
        if np_dt is None:
            raise TypeError(f"unsupported tensor dtype {t.dtype}")
        return t.data_ptr(), t.is_cuda, tuple(t.shape), np.dtype(np_dt), (t.device.index if t.is_cuda else None), t
    if hasattr(obj, "__dlpack__"):  # any other DLPack producer (jax, cupy without the CUDA array interface, numpy-likes, ...)
        return _unwrap_dlpack(obj, want_write)
    raise TypeError(f"unsupported buffer type {type(obj)!r}")


# ---- DLPack (https://dmlc.github.io/dlpack: struct DLManagedTensor, capsule name "dltensor") -------------------------------


class _DLDevice(C.Structure):
    _fields_ = [("device_type", C.c_int32), ("device_id", C.c_int32)]


class _DLDataType(C.Structure):
    _fields_ = [("code", C.c_uint8), ("bits", C.c_uint8), ("lanes", C.c_uint16)]


class _DLTensor(C.Structure):
    _fields_ = [("data", C.c_void_p), ("device", _DLDevice), ("ndim", C.c_int32), ("dtype", _DLDataType),
                ("shape", C.POINTER(C.c_int64)), ("strides", C.POINTER(C.c_int64)), ("byte_offset", C.c_uint64)]


class _DLManagedTensor(C.Structure):
    _fields_ = [("dl_tensor", _DLTensor), ("manager_ctx", C.c_void_p), ("deleter", C.c_void_p)]


_DL_CPU, _DL_CUDA, _DL_CUDA_HOST, _DL_CUDA_MANAGED = 1, 2, 3, 13
_DL_FLOAT = 2


def _unwrap_dlpack(obj, want_write=False, stream=None):
    """Read pointer, shape and dtype out of the DLPack capsule of `obj`.  The capsule is NOT consumed (not renamed): it is
    returned as the keep-alive object and its own destructor releases the producer's tensor when the call is over.  A CUDA
    producer is asked to make the data visible to `stream` (default: the legacy default stream, which is what the library
    orders against when no stream is given)."""
    dev_type, dev_id = (obj.__dlpack_device__() if hasattr(obj, "__dlpack_device__") else (_DL_CPU, 0))
    dev_type = int(dev_type)
    if dev_type in (_DL_CUDA, _DL_CUDA_MANAGED):
        cap = obj.__dlpack__(stream=1 if stream in (None, 0) else int(stream))
    else:
        cap = obj.__dlpack__()
    api = C.pythonapi
    api.PyCapsule_IsValid.argtypes, api.PyCapsule_IsValid.restype = [C.py_object, C.c_char_p], C.c_int
    api.PyCapsule_GetPointer.argtypes, api.PyCapsule_GetPointer.restype = [C.py_object, C.c_char_p], C.c_void_p
    if not api.PyCapsule_IsValid(cap, b"dltensor"):
        raise TypeError(f"{type(obj)!r}.__dlpack__() did not return an unconsumed 'dltensor' capsule")
    t = C.cast(api.PyCapsule_GetPointer(cap, b"dltensor"), C.POINTER(_DLManagedTensor)).contents.dl_tensor
    if t.device.device_type not in (_DL_CPU, _DL_CUDA, _DL_CUDA_HOST, _DL_CUDA_MANAGED):
        raise TypeError(f"DLPack device type {t.device.device_type} unsupported (CPU, CUDA, pinned or managed memory)")
    if t.dtype.code != _DL_FLOAT or t.dtype.lanes != 1 or t.dtype.bits not in (32, 64):
        raise TypeError(f"DLPack dtype (code {t.dtype.code}, {t.dtype.bits} bits, {t.dtype.lanes} lanes) unsupported (float32/float64)")
    shape = tuple(int(t.shape[i]) for i in range(t.ndim))
    if t.strides:  # in ELEMENTS; NULL = compact row-major
        acc = 1
        for i in reversed(range(t.ndim)):
            if shape[i] != 1 and int(t.strides[i]) != acc:
                raise ValueError("DLPack tensor must be C-contiguous")
            acc *= shape[i]
    on_dev = t.device.device_type in (_DL_CUDA, _DL_CUDA_MANAGED)
    ptr = int(t.data or 0) + int(t.byte_offset)
    return ptr, on_dev, shape, np.dtype(np.float32 if t.dtype.bits == 32 else np.float64), (int(t.device.device_id) if on_dev else None), cap


class Handle:
    """One library handle = one GPU.  Not thread-safe (serialise calls, like Keras predict)."""

    def __init__(self, device: int = 0):
        self._lib = load()
        self._h = C.c_void_p()
        _check(self._lib.vae21_create(int(device), C.byref(self._h)))
        self.device = int(device)
        self.dims: Optional[Tuple[int, ...]] = None
        self._fin = weakref.finalize(self, self._lib.vae21_destroy, self._h)

    def close(self):
        self._fin()

    # -- configuration
    def set_model(self, kernels: Sequence[np.ndarray], biases: Sequence[np.ndarray], relu: Sequence[bool]):
        n = len(kernels)
        ks = [np.ascontiguousarray(k, dtype=np.float32) for k in kernels]
        bs = [np.ascontiguousarray(b, dtype=np.float32) for b in biases]
        dims = [ks[0].shape[0]] + [k.shape[1] for k in ks]
        for i, (k, b) in enumerate(zip(ks, bs)):
            if k.shape != (dims[i], dims[i + 1]) or b.shape != (dims[i + 1],):
                raise ValueError(f"layer {i}: kernel {k.shape} / bias {b.shape} inconsistent")
        c_dims = (C.c_int * (n + 1))(*dims)
        c_k = (C.c_void_p * n)(*[k.ctypes.data for k in ks])
        c_b = (C.c_void_p * n)(*[b.ctypes.data for b in bs])
        c_r = (C.c_int * n)(*[1 if r else 0 for r in relu])
        _check(self._lib.vae21_set_model(self._h, n, c_dims, c_k, c_b, c_r))
        self.dims = tuple(dims)

    def set_norm(self, par_min, par_max, log_mask, floor_col, fx_floor, sig_mean, sig_std):
        pmin = np.ascontiguousarray(par_min, dtype=np.float64)
        pmax = np.ascontiguousarray(par_max, dtype=np.float64)
        mask = np.ascontiguousarray(log_mask, dtype=np.int32)
        mu = np.ascontiguousarray(sig_mean, dtype=np.float32)
        _check(self._lib.vae21_set_norm(
            self._h, len(pmin), pmin.ctypes.data_as(C.POINTER(C.c_double)), pmax.ctypes.data_as(C.POINTER(C.c_double)),
            mask.ctypes.data_as(C.POINTER(C.c_int)), int(floor_col), float(fx_floor), len(mu),
            mu.ctypes.data_as(C.POINTER(C.c_float)), float(sig_std)))

    # -- info
    def info(self):
        n, ms, tc = C.c_int64(0), C.c_float(0), C.c_int(0)
        _check(self._lib.vae21_get_info(self._h, C.byref(n), C.byref(ms), C.byref(tc)))
        return {"kernel_launches": n.value, "last_kernel_ms": ms.value, "tc_supported": bool(tc.value)}

    def tc_saturation(self, reset=False) -> int:
        """Epilogue threads that converted a hidden activation beyond the range of the fp16-based tensor-core operand formats
        since the last reset (vae21_get_tc_stats); 0 = inside the range the error budget was pinned for."""
        n = C.c_int64(0)
        _check(self._lib.vae21_get_tc_stats(self._h, C.byref(n), int(bool(reset))))
        return int(n.value)

    # -- compute
    def _prep_in(self, x, width):
        ptr, dev, shape, dt, didx, keep = _unwrap(x)
        if len(shape) != 2 or shape[1] != width:
            raise ValueError(f"expected shape (n, {width}), got {shape}")
        if dev and didx is not None and didx != self.device:
            raise ValueError(f"buffer on cuda:{didx}, handle on cuda:{self.device}")
        return ptr, dev, shape[0], dt, keep

    def _prep_out(self, out, n, width, device_like=None):
        if out is None:
            if device_like is not None:
                import torch

                out = torch.empty((n, width), dtype=torch.float32, device=device_like) if width else \
                    torch.empty((n,), dtype=torch.float32, device=device_like)
            else:
                out = pinned_empty((n, width) if width else (n,), np.float32)
        ptr, dev, shape, dt, didx, keep = _unwrap(out, want_write=True)
        exp = (n, width) if width else (n,)
        if tuple(shape) != exp or dt != np.float32:
            raise ValueError(f"output must be float32 with shape {exp}, got {dt} {shape}")
        return out, ptr, dev

    @staticmethod
    def _stream_ptr(stream, *buffers):
        """The CUDA stream the library must order its work against: the caller's `stream`, else the current torch stream of
        the first device-resident torch tensor among `buffers` (input OR output -- a device input produced on that stream must
        be complete before any library stream reads it), else the legacy default stream."""
        if stream is not None:
            return int(getattr(stream, "cuda_stream", stream))
        for b in buffers:
            if b is not None and getattr(b, "is_cuda", False):
                import torch

                return int(torch.cuda.current_stream(b.device).cuda_stream)
        return 0

    def predict(self, params, out=None, precision=FP32_SIMT, stream=None):
        if self.dims is None:
            raise Vae21Error(2, "model not set")
        ptr, dev, n, dt, keep = self._prep_in(params, self.dims[0])
        if dt not in (np.float32, np.float64):
            raise TypeError(f"params dtype {dt} unsupported (float32/float64)")
        out, optr, odev = self._prep_out(out, n, self.dims[-1], keep.device if dev and hasattr(keep, "is_cuda") else None)
        _check(self._lib.vae21_predict(self._h, ptr, F64 if dt == np.float64 else F32, int(dev), n, optr, int(odev),
                                       int(precision), self._stream_ptr(stream, keep, out)))
        return out

    def forward_normalised(self, x, out=None, precision=FP32_SIMT, stream=None):
        if self.dims is None:
            raise Vae21Error(2, "model not set")
        ptr, dev, n, dt, keep = self._prep_in(x, self.dims[0])
        if dt != np.float32:
            raise TypeError("normalised input must be float32")
        out, optr, odev = self._prep_out(out, n, self.dims[-1], keep.device if dev and hasattr(keep, "is_cuda") else None)
        _check(self._lib.vae21_forward_normalised(self._h, ptr, int(dev), n, optr, int(odev), int(precision),
                                                  self._stream_ptr(stream, keep, out)))
        return out

    def chi2(self, params, obs, inv_sigma, out=None, want_chi2=True, want_best=True, precision=FP32_SIMT, stream=None):
        """Returns (chi2 array or None, best_val, best_idx)."""
        if self.dims is None:
            raise Vae21Error(2, "model not set")
        ptr, dev, n, dt, keep = self._prep_in(params, self.dims[0])
        if dt not in (np.float32, np.float64):
            raise TypeError(f"params dtype {dt} unsupported (float32/float64)")
        nout = self.dims[-1]
        obs = np.ascontiguousarray(obs, dtype=np.float32)
        isg = np.ascontiguousarray(np.broadcast_to(np.asarray(inv_sigma, dtype=np.float32), (nout,)))
        if obs.shape != (nout,):
            raise ValueError(f"obs must have shape ({nout},)")
        optr, odev = None, int(dev)
        if want_chi2:
            out, optr, odev = self._prep_out(out, n, 0, keep.device if dev and hasattr(keep, "is_cuda") else None)
        else:
            out = None
        bv, bi = C.c_float(float("nan")), C.c_int64(-1)
        _check(self._lib.vae21_chi2(
            self._h, ptr, F64 if dt == np.float64 else F32, int(dev), n, obs.ctypes.data_as(C.POINTER(C.c_float)),
            isg.ctypes.data_as(C.POINTER(C.c_float)), optr, int(odev), C.byref(bv) if want_best else None,
            C.byref(bi) if want_best else None, int(precision), self._stream_ptr(stream, keep, out)))
        return out, (bv.value if want_best else None), (bi.value if want_best else None)

    def error(self, params, truth, band_mask=None, relative=True, precision=FP32_SIMT):
        """Fused emulator.py `error`: per-row rms difference between predict(params) and truth over the bins of `band_mask`
        (None = all), in % of the row's amplitude in the band (relative) or in mK.  Returns a float32 numpy array (n,)."""
        if self.dims is None:
            raise Vae21Error(2, "model not set")
        ptr, dev, n, dt, keep = self._prep_in(params, self.dims[0])
        if dt not in (np.float32, np.float64):
            raise TypeError(f"params dtype {dt} unsupported (float32/float64)")
        nout = self.dims[-1]
        tptr, tdev, tshape, tdt, _, tkeep = _unwrap(truth)
        if tuple(tshape) != (n, nout) or np.dtype(tdt) != np.float32:
            raise ValueError(f"truth must be float32 with shape ({n}, {nout})")
        mask = None
        if band_mask is not None:
            mask = np.ascontiguousarray(band_mask, dtype=np.float32)
            if mask.shape != (nout,):
                raise ValueError(f"band_mask must have shape ({nout},)")
        out = np.empty(n, np.float32)
        _check(self._lib.vae21_error(self._h, ptr, F64 if dt == np.float64 else F32, int(dev), n, tptr, int(tdev),
                                     mask.ctypes.data_as(C.POINTER(C.c_float)) if mask is not None else None, int(bool(relative)),
                                     out.ctypes.data, 0, int(precision), self._stream_ptr(None, keep, tkeep) or None))
        return out

    def chi2_grid(self, npts, obs, inv_sigma, x_lo=None, x_hi=None, first=0, count=None, out=None, precision=FP32_SIMT, stream=None):
        """Fused chi^2 over points [first, first + count) of a regular grid in normalised coordinates generated on the device
        (C order, last dimension fastest).  `out`: optional float32 DEVICE buffer of `count` elements.  Returns (best chi^2, global
        grid index of the best point)."""
        if self.dims is None:
            raise Vae21Error(2, "model not set")
        nd, nout = self.dims[0], self.dims[-1]
        npts = [int(v) for v in (npts if np.ndim(npts) else [npts] * nd)]
        if len(npts) != nd:
            raise ValueError(f"npts must have {nd} entries")
        total = int(np.prod(npts, dtype=np.int64))
        count = total - int(first) if count is None else int(count)
        lo = np.ascontiguousarray(np.broadcast_to(-1.0 if x_lo is None else np.asarray(x_lo, np.float64), (nd,)), dtype=np.float64)
        hi = np.ascontiguousarray(np.broadcast_to(1.0 if x_hi is None else np.asarray(x_hi, np.float64), (nd,)), dtype=np.float64)
        obs = np.ascontiguousarray(obs, dtype=np.float32)
        isg = np.ascontiguousarray(np.broadcast_to(np.asarray(inv_sigma, dtype=np.float32), (nout,)))
        if obs.shape != (nout,):
            raise ValueError(f"obs must have shape ({nout},)")
        optr, keep = None, None
        if out is not None:
            optr, dev, shape, dt, _, keep = _unwrap(out, want_write=True)
            if not dev or np.dtype(dt) != np.float32 or int(np.prod(shape)) < count:
                raise ValueError("out must be a float32 device buffer with at least `count` elements")
        bv, bi = C.c_float(float("nan")), C.c_int64(-1)
        _check(self._lib.vae21_chi2_grid(
            self._h, nd, (C.c_int32 * nd)(*npts), lo.ctypes.data_as(C.POINTER(C.c_double)), hi.ctypes.data_as(C.POINTER(C.c_double)),
            int(first), count, obs.ctypes.data_as(C.POINTER(C.c_float)), isg.ctypes.data_as(C.POINTER(C.c_float)), optr, C.byref(bv),
            C.byref(bi), int(precision), C.c_void_p(int(stream)) if stream else None))
        return bv.value, bi.value

    def mcmc_run(self, x_dev, logp_dev, lo, hi, obs, inv_sigma, a=2.0, seed=0, first_step=0, n_steps=1, init_logp=False,
                 precision=FP32_SIMT, stream=None, want_accepted=True):
        """n_steps stretch-move steps of the ensemble held in `x_dev` (float64 DEVICE, (walkers, n_par), coordinates of the
        prior box) / `logp_dev` (float64 DEVICE, (walkers,)), in place.  Returns the number of accepted proposals (None when
        `want_accepted` is false: the call then does not synchronise)."""
        if self.dims is None:
            raise Vae21Error(2, "model not set")
        nd, nout = self.dims[0], self.dims[-1]
        xptr, xdev, xshape, xdt, _, k1 = _unwrap(x_dev, want_write=True)
        lptr, ldev, lshape, ldt, _, k2 = _unwrap(logp_dev, want_write=True)
        if not (xdev and ldev) or np.dtype(xdt) != np.float64 or np.dtype(ldt) != np.float64:
            raise ValueError("x_dev / logp_dev must be float64 device arrays")
        if len(xshape) != 2 or xshape[1] != nd or tuple(lshape) != (xshape[0],):
            raise ValueError(f"x_dev must have shape (walkers, {nd}) and logp_dev (walkers,)")
        lo = np.ascontiguousarray(np.broadcast_to(np.asarray(lo, np.float64), (nd,)), dtype=np.float64)
        hi = np.ascontiguousarray(np.broadcast_to(np.asarray(hi, np.float64), (nd,)), dtype=np.float64)
        obs = np.ascontiguousarray(obs, dtype=np.float32)
        isg = np.ascontiguousarray(np.broadcast_to(np.asarray(inv_sigma, dtype=np.float32), (nout,)))
        if obs.shape != (nout,):
            raise ValueError(f"obs must have shape ({nout},)")
        acc = C.c_int64(-1)
        _check(self._lib.vae21_mcmc_run(
            self._h, xptr, lptr, int(xshape[0]), nd, lo.ctypes.data_as(C.POINTER(C.c_double)), hi.ctypes.data_as(C.POINTER(C.c_double)),
            obs.ctypes.data_as(C.POINTER(C.c_float)), isg.ctypes.data_as(C.POINTER(C.c_float)), float(a), int(seed) & (2**64 - 1),
            int(first_step), int(n_steps), 1 if init_logp else 0, int(precision), C.c_void_p(int(stream)) if stream else None,
            C.byref(acc) if want_accepted else None))
        return acc.value if want_accepted else None

    def time_predict(self, params_dev, out_dev, precision=FP32_SIMT, iters=10) -> float:
        ptr, dev, n, dt, keep = self._prep_in(params_dev, self.dims[0])
        out, optr, odev = self._prep_out(out_dev, n, self.dims[-1])
        if not (dev and odev):
            raise ValueError("time_predict needs device-resident buffers")
        ms = C.c_float(0)
        _check(self._lib.vae21_time_predict(self._h, ptr, F64 if dt == np.float64 else F32, n, optr, int(precision),
                                            int(iters), C.byref(ms)))
        return ms.value


# ---- trainer ---------------------------------------------------------------------------------


class Trainer:
    """ctypes face of `vae21_trainer` (include/vae21.h): fp32 parameters + Adam moments of a Dense stack on one GPU.
    Device buffers are passed as anything `_unwrap` understands (torch CUDA tensors, __cuda_array_interface__)."""

    def __init__(self, dims, relu, max_batch=256, device=0):
        self._lib = load()
        self.dims = [int(d) for d in dims]
        self.relu = [int(bool(r)) for r in relu]
        self.device = int(device)
        self.max_batch = int(max_batch)
        h = C.c_void_p()
        n = len(self.dims) - 1
        _check(self._lib.vae21_trainer_create(self.device, n, (C.c_int32 * (n + 1))(*self.dims), (C.c_int32 * n)(*self.relu),
                                              self.max_batch, C.byref(h)))
        self._t = h
        cnt = C.c_int64(0)
        _check(self._lib.vae21_trainer_num_params(self._t, C.byref(cnt)))
        self.num_params = int(cnt.value)
        self._fin = weakref.finalize(self, self._lib.vae21_trainer_destroy, h)

    def close(self):
        self._fin()

    def set_params(self, flat, reset_moments=True):
        flat = np.ascontiguousarray(flat, dtype=np.float32)
        if flat.size != self.num_params:
            raise ValueError(f"expected {self.num_params} parameters, got {flat.size}")
        _check(self._lib.vae21_trainer_set_params(self._t, flat.ctypes.data_as(C.POINTER(C.c_float)), int(bool(reset_moments))))

    def get_params(self) -> np.ndarray:
        out = np.empty(self.num_params, np.float32)
        _check(self._lib.vae21_trainer_get_params(self._t, out.ctypes.data_as(C.POINTER(C.c_float))))
        return out

    def set_moments(self, m, v):
        """Adam first / second moments, flat in parameter order (host arrays)."""
        m, v = np.ascontiguousarray(m, dtype=np.float32), np.ascontiguousarray(v, dtype=np.float32)
        if m.size != self.num_params or v.size != self.num_params:
            raise ValueError(f"expected {self.num_params} moments, got {m.size} / {v.size}")
        fp = C.POINTER(C.c_float)
        _check(self._lib.vae21_trainer_set_moments(self._t, m.ctypes.data_as(fp), v.ctypes.data_as(fp)))

    def get_moments(self):
        m, v = np.empty(self.num_params, np.float32), np.empty(self.num_params, np.float32)
        fp = C.POINTER(C.c_float)
        _check(self._lib.vae21_trainer_get_moments(self._t, m.ctypes.data_as(fp), v.ctypes.data_as(fp)))
        return m, v

    @staticmethod
    def _dev_ptr(obj, dtype, what):
        if obj is None:
            return None
        ptr, dev, _, dt, _, _ = _unwrap(obj)
        if not dev:
            raise ValueError(f"{what} must be a device buffer")
        if np.dtype(dt) != np.dtype(dtype):
            raise ValueError(f"{what} must be {np.dtype(dtype).name}, got {np.dtype(dt).name}")
        return ptr

    def forward_backward(self, x_all, y_all, w_all, batch, grad_scale, grad, loss_sum, idx=None, first=0, stream=None):
        """One batch: rows idx[0:batch] (int32 device array) or first..first+batch of the resident set.
        grad=None: forward + loss only.  loss_sum (1-element float32 device buffer) is incremented."""
        _check(self._lib.vae21_trainer_forward_backward(
            self._t, self._dev_ptr(x_all, np.float32, "x_all"), self._dev_ptr(y_all, np.float32, "y_all"),
            self._dev_ptr(w_all, np.float32, "w_all"), self._dev_ptr(idx, np.int32, "idx"), int(first), int(batch),
            float(grad_scale), self._dev_ptr(grad, np.float32, "grad"), self._dev_ptr(loss_sum, np.float32, "loss_sum"),
            C.c_void_p(int(stream)) if stream else None))

    def adam(self, grad, lr_t, beta1=0.9, beta2=0.999, eps=1e-7, stream=None):
        _check(self._lib.vae21_trainer_adam(self._t, self._dev_ptr(grad, np.float32, "grad"), float(lr_t), float(beta1),
                                            float(beta2), float(eps), C.c_void_p(int(stream)) if stream else None))

    def dp_begin(self, x_all, y_all, w_all, perm, n, batch, share_first, share_rows, lr, beta1, beta2, eps, iterations_before, grad,
                 loss_sum, stream=None):
        """Prepare a data-parallel epoch on this rank (see vae21_trainer_dp_begin): two graphs around the caller's all-reduce."""
        _check(self._lib.vae21_trainer_dp_begin(
            self._t, self._dev_ptr(x_all, np.float32, "x_all"), self._dev_ptr(y_all, np.float32, "y_all"),
            self._dev_ptr(w_all, np.float32, "w_all"), self._dev_ptr(perm, np.int32, "perm"), int(n), int(batch), int(share_first),
            int(share_rows), float(lr), float(beta1), float(beta2), float(eps), int(iterations_before),
            self._dev_ptr(grad, np.float32, "grad"), self._dev_ptr(loss_sum, np.float32, "loss_sum"),
            C.c_void_p(int(stream)) if stream else None))

    def dp_forward_backward(self, stream=None):
        _check(self._lib.vae21_trainer_dp_forward_backward(self._t, C.c_void_p(int(stream)) if stream else None))

    def dp_adam(self, stream=None):
        _check(self._lib.vae21_trainer_dp_adam(self._t, C.c_void_p(int(stream)) if stream else None))

    def epoch(self, x_all, y_all, w_all, perm, n, batch, lr, beta1, beta2, eps, iterations_before, loss_sum, stream=None):
        """All batches of one epoch in one library call (single GPU); see vae21_trainer_epoch."""
        _check(self._lib.vae21_trainer_epoch(
            self._t, self._dev_ptr(x_all, np.float32, "x_all"), self._dev_ptr(y_all, np.float32, "y_all"),
            self._dev_ptr(w_all, np.float32, "w_all"), self._dev_ptr(perm, np.int32, "perm"), int(n), int(batch), float(lr),
            float(beta1), float(beta2), float(eps), int(iterations_before), self._dev_ptr(loss_sum, np.float32, "loss_sum"),
            C.c_void_p(int(stream)) if stream else None))

    def launches(self) -> int:
        n = C.c_int64(0)
        _check(self._lib.vae21_trainer_launches(self._t, C.byref(n)))
        return int(n.value)
