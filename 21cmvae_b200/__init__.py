"""B200-native batched evaluation of the 21cmVAE global-signal emulator.

Drop-in for the hot path of christianhbye/21cmVAE
(``VeryAccurateEmulator.emulator.DirectEmulator.predict`` and the
``preprocess`` transforms it calls).  The arithmetic runs in one CUDA
library (``csrc/``, C-ABI in ``include/vae21.h``) loaded through ctypes;
there is no CPU fallback: ``predict`` raises if the library or a GPU is
missing.

The directory name is not a Python identifier; import it with
``importlib.import_module("21cmvae_b200")`` or through the top-level
``VeryAccurateEmulator`` shim package, which mirrors the reference's import
paths (``from VeryAccurateEmulator import emulator, preprocess``).
"""

__version__ = "0.1.0"

from . import preprocess  # noqa: F401  (pure numpy, no CUDA needed)
from . import keras_h5  # noqa: F401
