"""Row N3, first pass: `MyMSA.generate_msa` (savemoney/modules/msa.py:882-982) on the device.

The reference stacks the reads assigned to a plasmid -- each with the column CIGAR the alignment stage produced --
into one multiple alignment by appending one character per read and column in a Python loop ("time consuming" in
MyMSAligner.execute, msa.py:629).  Here the same layout is computed column-parallel by `smx_msa_run`
(csrc/smx_msa.cuh): extra columns per reference gap from the reads' I / S blocks, then one thread per read x gap.

    msa = generate_msa(ref_seq, query_seq_list, q_scores_list, my_cigar_list, param_dict)    # same argument order
    msa.ref_seq_aligned, msa.query_seq_list_aligned, msa.q_scores_list_aligned, msa.my_cigar_list_aligned

The four fields hold exactly what the reference's MyMSA holds after generate_msa; they are rendered from the
[n_reads, n_columns] byte matrices on first use (`*_matrix` gives the matrices themselves).  The Bayesian consensus
(msa.py:1746-1812, long double) over those matrices is consensus.py (`MyMSA.calculate_consensus`).
"""
from __future__ import annotations

import ctypes

import numpy as np

from .engine import EngineError

_OPS = {"I": 1, "D": 2, "S": 4, "H": 5, "=": 7, "X": 8}
_OP_OF = np.zeros(256, dtype=np.uint32)
for _c, _v in _OPS.items():
    _OP_OF[ord(_c)] = _v


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def cigar_runs(my_cigar: str) -> np.ndarray:
    """column string -> uint32 runs (len << 4 | op); letters outside = X I D S H are rejected like the reference does
    (`raise Exception(f"error!: {L}, {i}")`, msa.py:927)"""
    if not my_cigar:
        return np.zeros(0, np.uint32)
    b = np.frombuffer(my_cigar.encode("ascii"), dtype=np.uint8)
    ops = _OP_OF[b]
    if (ops == 0).any():
        bad = chr(int(b[np.nonzero(ops == 0)[0][0]]))
        raise Exception(f"error!: {bad}")
    cut = np.flatnonzero(np.diff(b)) + 1
    starts = np.concatenate(([0], cut))
    lens = np.diff(np.concatenate((starts, [b.size]))).astype(np.uint32)
    return (lens << 4) | ops[starts]


class MyMSA:
    """what generate_msa returns: the four aligned fields of the reference's MyMSA (msa.py:787-790)"""

    def __init__(self, ref_aligned, seq, qs, cig, param_dict):
        self.ref_seq_aligned_matrix, self.query_seq_matrix, self.q_scores_matrix, self.my_cigar_matrix = ref_aligned, seq, qs, cig
        self.param_dict = param_dict
        self._cache = {}

    sbq_pdf = None   # class attribute as in the reference (set by MyMSA.set_sbq_pdf, msa.py:1738-1745): a consensus.SequenceBasecallQscorePDF

    def calculate_consensus(self, engine=None):
        """msa.py:1749-1754 on the device (consensus.py): fills with_prior_consensus_* / without_prior_consensus_*"""
        from .consensus import calculate_consensus
        return calculate_consensus(self, type(self).sbq_pdf, engine=engine)

    @property
    def n_columns(self):
        return int(self.ref_seq_aligned_matrix.size)

    @property
    def ref_seq_aligned(self) -> str:
        return self.ref_seq_aligned_matrix.tobytes().decode("ascii")

    def _rows(self, key, mat):
        if key not in self._cache:
            self._cache[key] = [row.tobytes().decode("ascii") for row in mat]
        return self._cache[key]

    @property
    def query_seq_list_aligned(self):
        return self._rows("seq", self.query_seq_matrix)

    @property
    def my_cigar_list_aligned(self):
        return self._rows("cig", self.my_cigar_matrix)

    @property
    def q_scores_list_aligned(self):
        if "qs" not in self._cache:
            self._cache["qs"] = [row.tolist() for row in self.q_scores_matrix]
        return self._cache["qs"]


def generate_msa(ref_seq, query_seq_list, q_scores_list, my_cigar_list, param_dict=None, engine=None) -> MyMSA:
    """`my_cigar_list` entries are column strings (the reference's my_cigar) or uint32 run arrays as the engine delivers them."""
    from .savemoney import _seq_of, default_engine
    eng = engine or default_engine()
    n = len(query_seq_list)
    if not (len(q_scores_list) == len(my_cigar_list) == n):
        raise AssertionError("query_seq_list, q_scores_list and my_cigar_list differ in length")   # msa.py:888
    ref = np.frombuffer(_seq_of(ref_seq).encode("ascii"), dtype=np.uint8)
    qlen = np.fromiter((len(q) for q in query_seq_list), dtype=np.int32, count=n)
    qoff = np.zeros(n, np.int64)
    if n > 1:
        qoff[1:] = np.cumsum(qlen[:-1], dtype=np.int64)
    qry = np.frombuffer("".join(str(q) for q in query_seq_list).encode("ascii"), dtype=np.uint8) if n else np.zeros(1, np.uint8)
    for q, s in zip(query_seq_list, q_scores_list):
        if len(s) != len(q):
            raise AssertionError("a query and its q_scores differ in length")
    qs = np.concatenate([np.asarray(s, dtype=np.int8) for s in q_scores_list]) if n and int(qlen.sum()) else np.zeros(1, np.int8)
    runs = [c if isinstance(c, np.ndarray) else cigar_runs(c) for c in my_cigar_list]
    coff = np.zeros(n + 1, np.int64)
    if n:
        coff[1:] = np.cumsum([r.size for r in runs])
    rle = np.ascontiguousarray(np.concatenate(runs), dtype=np.uint32) if n and int(coff[-1]) else np.zeros(1, np.uint32)
    qry = np.ascontiguousarray(qry if qry.size else np.zeros(1, np.uint8))
    qs = np.ascontiguousarray(qs)
    ncols = ctypes.c_int64(0)
    lib = eng._lib
    rc = lib.smx_msa_run(eng._h, _ptr(ref) if ref.size else _ptr(np.zeros(1, np.uint8)), int(ref.size), n, _ptr(qry), _ptr(qs), _ptr(qoff if n else np.zeros(1, np.int64)),
                         _ptr(qlen if n else np.zeros(1, np.int32)), _ptr(rle), _ptr(coff), ctypes.byref(ncols))
    if rc == -6:
        raise AssertionError(lib.smx_last_error(eng._h).decode())   # the reference's final assertions (msa.py:913-916)
    eng._check(rc, "smx_msa_run")
    c = int(ncols.value)
    ref_al = np.empty(max(c, 1), np.uint8)
    seq = np.empty((n, c), np.uint8); cig = np.empty((n, c), np.uint8); q = np.empty((n, c), np.int8)
    eng._check(lib.smx_msa_fetch(eng._h, _ptr(ref_al), _ptr(seq) if seq.size else None, _ptr(q) if q.size else None, _ptr(cig) if cig.size else None), "smx_msa_fetch")
    return MyMSA(ref_al[:c], seq, q, cig, param_dict)
