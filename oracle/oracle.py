"""ctypes front-end of the CPU oracle (oracle/parasail_port.c).

TEST INFRASTRUCTURE ONLY: may be imported from tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  The product package
(multiplexnanopore_b200/) must never import this module.

It also provides `parasail_shim()`, a module object exposing the slice of the `parasail`
Python API that the reference calls, backed by the C restatement, so the
unmodified reference Python can be driven in this container (oracle/ref_harness.py).
"""
from __future__ import annotations

import ctypes
import os
import subprocess
import types
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
MODE_NW, MODE_SG_QE_DE, MODE_SG_QB_DB, MODE_SW = 0, 1, 2, 3


class _PPResult(ctypes.Structure):
    _fields_ = [
        ("score", ctypes.c_int32),
        ("end_query", ctypes.c_int32),
        ("end_ref", ctypes.c_int32),
        ("beg_query", ctypes.c_int32),
        ("beg_ref", ctypes.c_int32),
        ("n_ops", ctypes.c_int64),
    ]


_lib = None


def build(force: bool = False) -> Path:
    so = HERE / "liboracle.so"
    srcs = [HERE / "parasail_port.c", HERE / "consensus_port.c"]
    if force or not so.exists() or any(src.exists() and so.stat().st_mtime < src.stat().st_mtime for src in srcs):
        subprocess.check_call(["make", "-C", str(HERE), str(so)], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(str(build()))
        _lib.pp_align.restype = ctypes.c_int
        _lib.pp_align.argtypes = [ctypes.c_int, ctypes.c_char_p, ctypes.c_int, ctypes.c_char_p, ctypes.c_int,
                                  ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                  ctypes.POINTER(_PPResult), ctypes.c_char_p]
        _lib.pp_score_only.restype = ctypes.c_int32
        _lib.pp_score_only.argtypes = [ctypes.c_int, ctypes.c_char_p, ctypes.c_int, ctypes.c_char_p, ctypes.c_int,
                                       ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int]
        _lib.pp_kmer_offset_scan.restype = None
        _lib.pp_kmer_offset_scan.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p]
        _lib.pp_matrix_create.restype = None
        _lib.pp_matrix_create.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
        _lib.pp_set_tie_rules.restype = None
        _lib.pp_set_tie_rules.argtypes = [ctypes.c_int] * 4
        _lib.pp_consensus_error_rate.restype = ctypes.c_int
        _lib.pp_consensus_error_rate.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_char_p, ctypes.c_void_p, ctypes.c_void_p,
                                                 ctypes.c_void_p, ctypes.c_void_p]
    return _lib


def consensus_error_rate(P25: np.ndarray, pdf: np.ndarray, called: str, q_scores, PN5) -> list:
    """SequenceBasecallQscorePDF.calc_consensus_error_rate (alignment_functions.pyx:93-132) for one column: p_list over the
    bases "ATCG-".  P25 [25] and pdf [25, 93] float64 in the key order B_c with B, c over "ATCG-"."""
    P25 = np.ascontiguousarray(P25, dtype=np.float64); pdf = np.ascontiguousarray(pdf, dtype=np.float64)
    q = np.ascontiguousarray(q_scores, dtype=np.int32)
    pn = np.ascontiguousarray(PN5, dtype=np.float64)
    out = np.zeros(5, np.float64)
    scratch = np.zeros(max(1, len(called)), np.longdouble)
    rc = lib().pp_consensus_error_rate(P25.ctypes.data, pdf.ctypes.data, len(called), called.encode("ascii"), q.ctypes.data, pn.ctypes.data,
                                       out.ctypes.data, scratch.ctypes.data)
    if rc != 0:
        raise ValueError("a character outside ATCG- or a q-score above 92 (the reference indexes out of bounds there)")
    return [float(x) for x in out]


def set_tie_rules(gap_prefers_extend: int = 1, h_order_diag_f_e: int = 1, sg_end_order: int = 0, sw_end_order: int = 1):
    """Flip the tie rules of parasail_port.c that could not be checked against a real parasail (defaults = what the
    product implements).  Only oracle/tie_sensitivity.py and its test call this with non-default values."""
    lib().pp_set_tie_rules(int(gap_prefers_extend), int(h_order_diag_f_e), int(sg_end_order), int(sw_end_order))


def rle(ops: str) -> str:
    """expanded column string -> '12=1X3I' (same text parasail's cigar.decode carries)."""
    if not ops:
        return ""
    out = []
    prev, n = ops[0], 0
    for c in ops:
        if c == prev:
            n += 1
        else:
            out.append(f"{n}{prev}")
            prev, n = c, 1
    out.append(f"{n}{prev}")
    return "".join(out)


class OracleResult:
    __slots__ = ("score", "end_query", "end_ref", "beg_query", "beg_ref", "ops")

    def __init__(self, r: _PPResult, ops: str):
        self.score, self.end_query, self.end_ref = r.score, r.end_query, r.end_ref
        self.beg_query, self.beg_ref, self.ops = r.beg_query, r.beg_ref, ops

    @property
    def cigar_rle(self) -> str:
        return rle(self.ops)


def align(mode: int, s1: str, s2: str, open_: int, gap: int, match: int, mismatch: int) -> OracleResult:
    """s1 = query, s2 = reference (parasail argument order)."""
    b1, b2 = s1.encode("ascii"), s2.encode("ascii")
    buf = ctypes.create_string_buffer(len(b1) + len(b2) + 2)
    r = _PPResult()
    rc = lib().pp_align(mode, b1, len(b1), b2, len(b2), open_, gap, match, mismatch, ctypes.byref(r), buf)
    if rc != 0:
        raise ValueError("oracle pp_align rejected its arguments (empty sequence?)")
    return OracleResult(r, buf.raw[: r.n_ops].decode("ascii"))


def score_only(mode: int, s1: str, s2: str, open_: int, gap: int, match: int, mismatch: int) -> int:
    b1, b2 = s1.encode("ascii"), s2.encode("ascii")
    return int(lib().pp_score_only(mode, b1, len(b1), b2, len(b2), open_, gap, match, mismatch))


def kmer_offset_scan(ref_rep: np.ndarray, query: np.ndarray, max_k: int, n_query: int) -> np.ndarray:
    """restates af.k_mer_offset_analysis_2 (alignment_functions.pyx:46-64)."""
    ref_rep = np.ascontiguousarray(ref_rep, dtype=np.int64)
    query = np.ascontiguousarray(query, dtype=np.int64)
    assert ref_rep.size >= max_k + n_query - 1 and query.size >= n_query
    out = np.empty(max_k, dtype=np.int64)
    lib().pp_kmer_offset_scan(ref_rep.ctypes.data, query.ctypes.data, max_k, n_query, out.ctypes.data)
    return out


# ------------------------------------------------------------------------------------------------
# `parasail`-shaped shim (only what ref_query_alignment.py / post_analysis_core.py touch)
# ------------------------------------------------------------------------------------------------
class _ShimMatrix:
    def __init__(self, alphabet: str, match: int, mismatch: int):
        assert alphabet == "ACGT", "the reference only ever builds an ACGT matrix (ref_query_alignment.py:32)"
        self.match, self.mismatch = match, mismatch
        m = np.zeros(25, dtype=np.int32)
        lib().pp_matrix_create(match, mismatch, m.ctypes.data)
        self.matrix = m.reshape(5, 5)
        self.size = 5


class _ShimCigar:
    def __init__(self, r: OracleResult):
        self.decode = r.cigar_rle.encode("ascii")
        self.beg_query, self.beg_ref = r.beg_query, r.beg_ref


class _ShimResult:
    def __init__(self, r: OracleResult):
        self.score, self.end_query, self.end_ref = r.score, r.end_query, r.end_ref
        self.cigar = _ShimCigar(r)


def parasail_shim(call_log: list | None = None) -> types.ModuleType:
    """module object to be registered as sys.modules['parasail'].

    call_log, when given, receives one (mode, len_s1, len_s2) tuple per DP call: the harvested
    problem list used for GCUPS accounting (SURVEY.md section 8d 'DP cell').
    """
    mod = types.ModuleType("parasail")

    def _make(mode):
        def f(s1, s2, open_, extend, matrix):
            if call_log is not None:
                call_log.append((mode, len(s1), len(s2)))
            return _ShimResult(align(mode, str(s1), str(s2), open_, extend, matrix.match, matrix.mismatch))
        return f

    mod.matrix_create = lambda alphabet, match, mismatch: _ShimMatrix(alphabet, match, mismatch)
    mod.nw_trace = _make(MODE_NW)
    mod.sg_qe_de_trace = _make(MODE_SG_QE_DE)
    mod.sg_qb_db_trace = _make(MODE_SG_QB_DB)
    mod.sw_trace = _make(MODE_SW)
    return mod
