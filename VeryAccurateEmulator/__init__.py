"""Import shim: ``from VeryAccurateEmulator import emulator, preprocess`` resolves to the
B200-native package in ``21cmvae_b200/`` (whose directory name is not a Python identifier),
so notebooks, MCMC drivers and parameter-grid scripts written against christianhbye/21cmVAE
keep their import lines.  Unlike the reference's ``__init__`` nothing is downloaded here."""
import importlib as _il
import os as _os
import sys as _sys

_root = _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__)))
if _root not in _sys.path:
    _sys.path.insert(0, _root)
_pkg = _il.import_module("21cmvae_b200")
__version__ = _pkg.__version__
for _name in ("preprocess", "emulator", "keras_h5", "multigpu", "training", "mcmc"):
    _mod = _il.import_module("21cmvae_b200." + _name)
    _sys.modules[__name__ + "." + _name] = _mod
    globals()[_name] = _mod
