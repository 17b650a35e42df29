"""Row N1: the re-alignment of read chunks against a chunk consensus that the reference's polish
stage runs one parasail call at a time -- `MyMSA.exec_chunk_poa` (savemoney/modules/msa.py:1249-1355), four polish
rounds per plasmid (msa.py:631-638) -- collected into ONE batch per round on the CUDA engine.

  BatchedAligner     the five alignment operations of MyAlignerBase (savemoney/modules/ref_query_alignment.py:21-163:
                     nw_trace, my_special_sg_qe_de, my_special_sg_qb_db, my_special_sw, my_special_DP) over lists of
                     calls: `smx_ops_run` plans the DP problems, runs the batched kernels (re-clipping on the device) and
                     joins the halves of my_special_DP on host threads
  ChunkRealigner     the per-read decisions of exec_chunk_poa (which operation, where the read's physical end cuts the
                     chunk) for every chunk of a round, and the batch behind them; the POA consensus itself (spoa) and
                     the column re-layout (generate_msa, row N3) stay with the caller

Same names / argument meaning as the reference; every result is the reference's `my_cigar` string.
"""
from __future__ import annotations

import ctypes
import re
from typing import Sequence

import numpy as np

from .engine import EngineError

NW, SG_QE_DE, SG_QB_DB, SW, SPECIAL_DP = 0, 1, 2, 3, 4
_LUT = np.zeros(16, dtype="S1")
for _k, _v in {1: "I", 2: "D", 4: "S", 5: "H", 7: "=", 8: "X"}.items():
    _LUT[_k] = _v.encode()


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


class BatchedAligner:
    """MyAlignerBase for many calls at once.

        al = BatchedAligner(param_dict, engine)
        cigars = al.run([(NW, ref, query), (SPECIAL_DP, ref, q1, q2), (SW, ref, query), ...])   # list of my_cigar strings
    """

    def __init__(self, param_dict, engine=None, n_threads=0):
        from .savemoney import default_engine
        self.go, self.ge = param_dict["gap_open_penalty"], param_dict["gap_extend_penalty"]
        self.ma, self.mi = param_dict["match_score"], param_dict["mismatch_score"]
        self.engine = engine or default_engine()
        self.n_threads = n_threads
        self.last_cells = 0
        self.last_ms = 0.0

    def run_arrays(self, ops: Sequence[tuple]):
        """-> (cigar_off int64[n+1], cigar_rle uint32[total], status uint8[n]); sequences are pooled once each"""
        eng = self.engine
        n = len(ops)
        index, seqs = {}, []

        def slot(s):
            k = index.get(s)
            if k is None:
                k = index[s] = len(seqs)
                seqs.append(s)
            return k

        kind = np.empty(n, np.uint8)
        sel = np.zeros((n, 3), np.int64)
        has = np.zeros((n, 3), bool)
        for i, op in enumerate(ops):
            kind[i] = op[0]
            for j, s in enumerate(op[1:4]):
                if s:
                    sel[i, j] = slot(s); has[i, j] = True
            if not has[i, 0]:
                raise ValueError(f"operation {i}: empty reference")
        if n == 0:
            return np.zeros(1, np.int64), np.zeros(0, np.uint32), np.zeros(0, np.uint8)
        offs, lens = eng.upload_sequences(seqs)
        ro, rl = offs[sel[:, 0]], lens[sel[:, 0]]
        q1o, q1l = np.where(has[:, 1], offs[sel[:, 1]], 0), np.where(has[:, 1], lens[sel[:, 1]], 0).astype(np.int32)
        q2o, q2l = np.where(has[:, 2], offs[sel[:, 2]], 0), np.where(has[:, 2], lens[sel[:, 2]], 0).astype(np.int32)
        arrs = [np.ascontiguousarray(a) for a in (kind, ro.astype(np.int64), rl.astype(np.int32), q1o.astype(np.int64), q1l, q2o.astype(np.int64), q2l)]
        total = ctypes.c_int64(0)
        eng._check(eng._lib.smx_ops_run(eng._h, *[_ptr(a) for a in arrs], n, self.go, self.ge, self.ma, self.mi, int(self.n_threads),
                                        ctypes.byref(total)), "smx_ops_run")
        off = np.zeros(n + 1, np.int64)
        rle = np.empty(max(int(total.value), 1), np.uint32)
        status = np.zeros(n, np.uint8)
        eng._check(eng._lib.smx_ops_fetch(eng._h, _ptr(off), _ptr(rle), rle.size, _ptr(status)), "smx_ops_fetch")
        self.last_cells = eng.align_cells
        self.last_ms = eng.last_run_ms
        return off, rle[: int(total.value)], status

    def run(self, ops: Sequence[tuple]) -> list:
        off, rle, status = self.run_arrays(ops)
        if status.any():
            bad = np.nonzero(status)[0]
            raise EngineError(f"operation(s) {bad[:8].tolist()}: the reference raises on this input (empty local alignment / failed assertion)")
        return expand_all(off, rle)

    # the reference's per-call names (each is a batch of one; use run() for throughput)
    def nw_trace(self, ref_seq, query_seq):
        return self.run([(NW, str(ref_seq), str(query_seq))])[0]

    def my_special_sg_qe_de(self, ref_seq, query_seq):
        return self.run([(SG_QE_DE, str(ref_seq), str(query_seq))])[0]

    def my_special_sg_qb_db(self, ref_seq, query_seq):
        return self.run([(SG_QB_DB, str(ref_seq), str(query_seq))])[0]

    def my_special_sw(self, ref_seq, query_seq):
        return self.run([(SW, str(ref_seq), str(query_seq))])[0]

    def my_special_DP(self, query_seq_1, query_seq_2, ref_seq):
        return self.run([(SPECIAL_DP, str(ref_seq), str(query_seq_1), str(query_seq_2))])[0]


def expand_all(off, rle) -> list:
    """run arrays -> one column string per operation (one np.repeat over everything)"""
    if rle.size == 0:
        return [""] * (len(off) - 1)
    cols = np.repeat(_LUT[rle & 15], rle >> 4).tobytes().decode("ascii")
    ends = np.zeros(len(off), np.int64)
    np.cumsum(rle >> 4, out=(csum := np.empty(rle.size, np.int64)))
    ends[1:] = np.where(off[1:] > 0, csum[np.maximum(off[1:], 1) - 1], 0)
    return [cols[ends[i]: ends[i + 1]] for i in range(len(off) - 1)]


# ------------------------------------------------------------------------------------------------------------------
# exec_chunk_poa: which operation every read chunk needs
# ------------------------------------------------------------------------------------------------------------------
_ONLY_HO = re.compile(r"[HO]+")
_ONLY_ND = re.compile(r"[ND]+")
_LEADING_S = re.compile(r"N*S")
_TRAILING_S = re.compile(r"SN*$")


def _bases(aligned: str) -> str:
    return aligned.replace(" ", "").replace("-", "")


def plan_read_chunk(consensus: str, query_chunk_aligned: str, cigar_chunk_aligned: str, cut: int):
    """One read's chunk -> ("literal", my_cigar) or an operation tuple for BatchedAligner.run.

    `cut` = position inside the aligned chunk at which the read's physical start lies (the reference computes it as
    len(ref_seq_aligned) - query_seq_offset_list_aligned[idx] - s_idx, wrapped once; msa.py:1285-1287, :1322-1324):
    when it falls inside the chunk the two sides are independent pieces of the read and go through my_special_DP.
    Decision order of msa.py:1251-1262 and :1283-1343."""
    n = len(consensus)
    if _ONLY_HO.fullmatch(cigar_chunk_aligned):
        return ("literal", "H" * n)
    if _ONLY_ND.fullmatch(cigar_chunk_aligned):
        return ("literal", "D" * n)
    inside = 0 <= cut < len(query_chunk_aligned)
    halves = (_bases(query_chunk_aligned[:cut]), _bases(query_chunk_aligned[cut:])) if inside else None
    query = _bases(query_chunk_aligned)
    if not any(c in cigar_chunk_aligned for c in "SHO"):          # "valid4poa"
        return (SPECIAL_DP, consensus, *halves) if inside else (NW, consensus, query)
    open_l, open_r = cigar_chunk_aligned[0] in "OH", cigar_chunk_aligned[-1] in "OH"
    if open_l and open_r:
        return (SW, consensus, query)
    if open_l:
        return (SG_QB_DB, consensus, query)
    if open_r:
        return (SG_QE_DE, consensus, query)
    s_l, s_r = _LEADING_S.match(cigar_chunk_aligned) is not None, _TRAILING_S.search(cigar_chunk_aligned) is not None
    if s_l and s_r:
        raise Exception("error")                                     # msa.py:1316-1317
    if inside:
        return (SPECIAL_DP, consensus, *halves)
    if s_l:
        return (SG_QB_DB, consensus, query)
    if s_r:
        return (SG_QE_DE, consensus, query)
    raise Exception("error")                                         # msa.py:1342-1343


class ChunkRealigner:
    """The DP part of one polish round for all chunks at once.

        cr = ChunkRealigner(param_dict, engine)
        my_cigar_lists = cr.realign(len_ref_seq_aligned, query_seq_offset_list_aligned, chunks)
            chunks: iterable of (consensus_wo_appendix, query_seq_chunk_list_aligned, my_cigar_chunk_list_aligned, s_idx)
        -> per chunk the `my_cigar_list` that exec_chunk_poa hands to generate_msa (msa.py:1346-1352)
    """

    def __init__(self, param_dict, engine=None, n_threads=0):
        self.aligner = BatchedAligner(param_dict, engine, n_threads)
        self.n_ops = 0

    def realign(self, len_ref_seq_aligned: int, query_seq_offset_list_aligned: Sequence[int], chunks) -> list:
        plans, ops = [], []
        for consensus, q_aligned_list, cigar_aligned_list, s_idx in chunks:
            row = []
            for idx, (qa, ca) in enumerate(zip(q_aligned_list, cigar_aligned_list)):
                cut = len_ref_seq_aligned - query_seq_offset_list_aligned[idx] - s_idx
                if cut < 0:
                    cut += len_ref_seq_aligned
                p = plan_read_chunk(consensus, qa, ca, cut)
                if p[0] == "literal":
                    row.append(p[1])
                else:
                    row.append(len(ops)); ops.append(p)
            plans.append(row)
        self.n_ops = len(ops)
        cigars = self.aligner.run(ops) if ops else []
        return [[cigars[x] if isinstance(x, int) else x for x in row] for row in plans]
