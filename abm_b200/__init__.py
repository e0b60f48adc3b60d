"""abm_b200 -- B200-native engine for the data-parallel hot path of scioip34/ABM:
per-agent visual-field projection + vision-driven movement / flocking update, batched
over agents and replicate simulations.  The arithmetic lives in hand-written sm_100a
CUDA kernels behind the C ABI of include/abm_b200.h (libabm_b200.so, loaded with
ctypes); this package is the host-side mirror of the reference's Python interface.
There is no CPU fallback."""
from ._lib import AbmError, LIB_PATH  # noqa: F401
from .engine import VFEngine  # noqa: F401
from .base_engine import BaseEngine  # noqa: F401

__version__ = "0.1.0"
