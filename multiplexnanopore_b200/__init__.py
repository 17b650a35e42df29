"""B200-native (sm_100a) engine for SAVEMONEY's read x plasmid affine-gap alignment path.

Only what the hot path needs lives here:
  csrc/      hand-written CUDA kernels + the C-ABI (include/smx.h)
  _lib.py    ctypes binding (no fallback)
  engine.py  batched host API over the C-ABI
"""
from . import _lib  # noqa: F401  (fails loudly when libsmx.so is absent)
from .engine import Engine, EngineError, AlignBatchResult  # noqa: F401

__version__ = "0.1.0"
