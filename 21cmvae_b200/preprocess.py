"""Host-side mirror of ``VeryAccurateEmulator.preprocess`` (same names, same
argument meaning, same dtypes) plus the cached-statistics form the CUDA
library consumes.

Reference: /root/reference/VeryAccurateEmulator/preprocess.py
  preproc        :4-24    (signal - mean(train, axis=0)) / std(train)
  unpreproc      :27-46   signal * std(train) + mean(train, axis=0)
  par_transform  :49-110  log10 of columns 0..2 (fx == 0 -> 1e-6), min/max map to [-1, 1]

The reference recomputes the training-set statistics on every call (about
50 ms, independent of batch size).  ``NormStats`` computes the same 466
numbers once; the kernels apply them in their prologue/epilogue, and these
numpy functions remain for callers that use the module directly (training
targets, tests, notebooks).
"""

from __future__ import annotations

from dataclasses import dataclass

import numpy as np

LOG_COLUMNS = (0, 1, 2)  # fstar, Vc, fx
FX_FLOOR = 10 ** (-6)


def _log_columns(p: np.ndarray) -> np.ndarray:
    """float64 copy of ``p`` with log10 applied to the first three columns."""
    head = p[:, :2].copy()
    fx = p[:, 2].copy()
    fx[fx == 0] = FX_FLOOR  # in the input's own dtype, like the reference
    t = np.empty(p.shape)
    with np.errstate(divide="ignore", invalid="ignore"):
        t[:, :2] = np.log10(head)
        t[:, 2] = np.log10(fx)
    t[:, 3:] = p[:, 3:]
    return t


@dataclass(frozen=True)
class NormStats:
    """The constants of the hot path, computed once from the training set."""

    par_min: np.ndarray  # (P,) float64, min over rows of the log-transformed training parameters
    par_max: np.ndarray  # (P,) float64
    sig_mean: np.ndarray  # (S,) signal dtype (float32 for the 21cmVAE dataset)
    sig_std: np.floating  # scalar, std over ALL training-signal elements (ddof = 0)

    @classmethod
    def from_training_set(cls, par_train: np.ndarray, signal_train: np.ndarray) -> "NormStats":
        t = _log_columns(np.asarray(par_train))
        signal_train = np.asarray(signal_train)
        return cls(np.min(t, axis=0), np.max(t, axis=0), np.mean(signal_train, axis=0), np.std(signal_train))


def preproc(signal: np.ndarray, signal_train: np.ndarray) -> np.ndarray:
    """Centre by the per-bin training mean, scale by the global training std."""
    out = signal.copy()
    out -= np.mean(signal_train, axis=0)
    out /= np.std(signal_train)
    return out


def unpreproc(signal: np.ndarray, signal_train: np.ndarray) -> np.ndarray:
    """Inverse of :func:`preproc` (multiply, then add -- two roundings)."""
    out = signal * np.std(signal_train)
    out += np.mean(signal_train, axis=0)
    return out


def par_transform(parameters: np.ndarray, params_train: np.ndarray) -> np.ndarray:
    """Map parameters with the affine map that sends the (log-transformed)
    training set onto [-1, 1] per column.  1-D input is treated as one row.
    Values outside the training range are not clipped.  Returns float64."""
    if len(np.shape(parameters)) == 1:
        parameters = np.expand_dims(parameters, axis=0)
    t = _log_columns(parameters)
    t_train = _log_columns(params_train)
    hi = np.max(t_train, axis=0)
    lo = np.min(t_train, axis=0)
    t -= lo
    t /= hi - lo
    t *= 2
    t -= 1
    return t


def par_transform_stats(parameters: np.ndarray, stats: NormStats) -> np.ndarray:
    """:func:`par_transform` with pre-computed statistics (identical result)."""
    if len(np.shape(parameters)) == 1:
        parameters = np.expand_dims(parameters, axis=0)
    t = _log_columns(np.asarray(parameters))
    t -= stats.par_min
    t /= stats.par_max - stats.par_min
    t *= 2
    t -= 1
    return t
