"""ctypes binding of libsmx.so -- the only way the Python host reaches the CUDA engine.

There is deliberately no fallback: if the shared library is missing or cannot be loaded the
import of the product package fails loudly (north_star: "no CPU fallback").
"""
from __future__ import annotations

import ctypes
from ctypes import POINTER, c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_void_p
from pathlib import Path

import os

# Six handles per GPU run five streams each (DP, high priority, three team streams); with the default of 8 hardware work
# queues streams share queues and kernels of different batches serialise behind each other (measured: +2 % end to end at
# 32).  Only effective when set before the process creates its CUDA context.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

_PKG = Path(__file__).resolve().parent
# SMX_LIB_PATH: developer override to A/B-test differently compiled builds of the same sources
LIB_PATH = Path(os.environ["SMX_LIB_PATH"]).resolve() if os.environ.get("SMX_LIB_PATH") else _PKG / "libsmx.so"

# every symbol include/smx.h declares: (restype, argtypes)
_SIGNATURES = {
    "smx_version": (c_int, []),
    "smx_device_count": (c_int, []),
    "smx_create": (c_int, [POINTER(c_void_p), c_int, c_void_p]),
    "smx_destroy": (None, [c_void_p]),
    "smx_last_error": (c_char_p, [c_void_p]),
    "smx_set_trace_budget": (c_int, [c_void_p, c_int64]),
    "smx_get_trace_budget": (c_int64, [c_void_p]),
    "smx_pool_upload": (c_int, [c_void_p, c_void_p, c_int64]),
    "smx_align_prepare": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int32,
                                  c_int32, c_int32, c_int32, c_int32]),
    "smx_align_run": (c_int, [c_void_p, POINTER(c_int64)]),
    "smx_align_fetch": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64]),
    "smx_align_kernel_ms": (c_int, [c_void_p, c_void_p, c_void_p]),
    "smx_debug_copy_trace": (c_int, [c_void_p, c_void_p, c_int64]),
    "smx_debug_guard_errors": (c_int64, [c_void_p]),
    "smx_last_run_ms": (c_float, [c_void_p]),
    "smx_align_cells": (c_int64, [c_void_p]),
    "smx_last_run_launches": (c_int32, [c_void_p]),
    "smx_scan_prepare": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_void_p]),
    "smx_scan_run": (c_int, [c_void_p]),
    "smx_scan_fetch": (c_int, [c_void_p, c_void_p, c_int64]),
    "smx_scan_cells": (c_int64, [c_void_p]),
    "smx_pipeline_run": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_void_p,
                                 c_void_p, c_void_p, c_void_p, c_int32, c_void_p, c_void_p, c_int32,
                                 c_int32, c_int32, c_int32, c_int32, c_int32, c_int32, c_int32, POINTER(c_int64)]),
    "smx_pipeline_fetch": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64]),
    "smx_pipeline_stats": (c_int, [c_void_p, POINTER(c_double), c_int32]),
    "smx_measure_int_peak": (c_int, [c_void_p, c_int32, POINTER(c_double)]),
    "smx_ops_run": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int32,
                            c_int32, c_int32, c_int32, c_int32, c_int32, POINTER(c_int64)]),
    "smx_ops_fetch": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
    "smx_msa_run": (c_int, [c_void_p, c_void_p, c_int32, c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, POINTER(c_int64)]),
    "smx_msa_fetch": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "smx_consensus_run": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_void_p, c_void_p, c_void_p]),
    "smx_assign": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_double, c_void_p, c_void_p, c_void_p]),
    "smx_ir_format": (c_int64, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int32, c_void_p, c_int64]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_lib = None


def load() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -m multiplexnanopore_b200.build` "
                "(or __graft_entry__.build()). The engine has no CPU fallback.")
        lib = ctypes.CDLL(str(LIB_PATH))
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib
