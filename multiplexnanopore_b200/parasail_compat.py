"""`parasail`-shaped per-call adapter over the batched CUDA engine ("per-call
compatibility surface").  Same names and argument meaning as the slice of parasail the reference uses:

    matrix_create("ACGT", match, mismatch)                      ref_query_alignment.py:32
    nw_trace / sg_qe_de_trace / sg_qb_db_trace / sw_trace       ref_query_alignment.py:35,58,68,82,343
        (s1 = query, s2 = reference, open, extend, matrix) -> Result with .score, .end_query, .end_ref,
        .cigar.decode (bytes, RLE), .cigar.beg_query, .cigar.beg_ref

One call = one problem = one kernel launch sequence, so this is for parity tests and for code that has not been
batched yet (e.g. msa.py's polish stage); throughput comes from `Engine.align_batch` / `savemoney.align_all`.
`align_many` takes a list of calls and runs them as ONE batch.
"""
from __future__ import annotations

import numpy as np

from .engine import Engine, MODE_NW, MODE_SG_QB_DB, MODE_SG_QE_DE, MODE_SW

_engine = None


def _eng() -> Engine:
    global _engine
    if _engine is None:
        _engine = Engine(0)
    return _engine


def use_engine(engine: Engine):
    global _engine
    _engine = engine


class Matrix:
    def __init__(self, alphabet: str, match: int, mismatch: int):
        if alphabet != "ACGT":
            raise ValueError("only the ACGT matrix of the reference is supported (ref_query_alignment.py:32)")
        self.match, self.mismatch, self.size = int(match), int(mismatch), 5
        m = np.full((5, 5), mismatch, dtype=np.int32)
        np.fill_diagonal(m, match)
        m[4, :] = 0
        m[:, 4] = 0
        self.matrix = m


def matrix_create(alphabet: str, match: int, mismatch: int) -> Matrix:
    return Matrix(alphabet, match, mismatch)


class Cigar:
    def __init__(self, decode: bytes, beg_query: int, beg_ref: int):
        self.decode, self.beg_query, self.beg_ref = decode, beg_query, beg_ref


class Result:
    def __init__(self, score, end_query, end_ref, cigar: Cigar):
        self.score, self.end_query, self.end_ref, self.cigar = score, end_query, end_ref, cigar


def align_many(calls, open_: int, extend: int, matrix: Matrix):
    """calls: iterable of (mode, s1, s2); one batched launch.  Returns a list of Result."""
    calls = list(calls)
    if any(len(s1) == 0 or len(s2) == 0 for _, s1, s2 in calls):
        raise ValueError("empty sequence")
    chunks, s1o, s1l, s2o, s2l, modes = [], [], [], [], [], []
    pos = 0
    for mode, s1, s2 in calls:
        a, b = str(s1).upper().encode("ascii"), str(s2).upper().encode("ascii")
        s1o.append(pos); s1l.append(len(a)); pos += len(a)
        s2o.append(pos); s2l.append(len(b)); pos += len(b)
        chunks += [a, b]; modes.append(mode)
    eng = _eng()
    eng.upload_pool(b"".join(chunks))
    r = eng.align_batch(s1o, s1l, s2o, s2l, modes, int(open_), int(extend), matrix.match, matrix.mismatch)
    return [Result(int(r.score[p]), int(r.end_query[p]), int(r.end_ref[p]),
                   Cigar(r.cigar_string(p).encode("ascii"), int(r.beg_query[p]), int(r.beg_ref[p]))) for p in range(len(calls))]


def _one(mode):
    def f(s1, s2, open_, extend, matrix):
        return align_many([(mode, s1, s2)], open_, extend, matrix)[0]
    return f


nw_trace = _one(MODE_NW)
sg_qe_de_trace = _one(MODE_SG_QE_DE)
sg_qb_db_trace = _one(MODE_SG_QB_DB)
sw_trace = _one(MODE_SW)
