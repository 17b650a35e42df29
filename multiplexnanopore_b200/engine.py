"""Batched host API over the C-ABI (include/smx.h).  numpy in, numpy out; no torch types."""
from __future__ import annotations

import ctypes
from dataclasses import dataclass
from typing import Sequence

import numpy as np

from . import _lib

MODE_NW, MODE_SG_QE_DE, MODE_SG_QB_DB, MODE_SW = 0, 1, 2, 3
TOPOLOGY_CIRCULAR, TOPOLOGY_LINEAR = 0, 1
OP_CHARS = {1: "I", 2: "D", 7: "=", 8: "X"}
_OP_LUT = np.zeros(16, dtype="S1")
for _k, _v in OP_CHARS.items():
    _OP_LUT[_k] = _v.encode()


class EngineError(RuntimeError):
    pass


def device_count() -> int:
    """CUDA devices visible to this process (smx_device_count)"""
    return int(_lib.load().smx_device_count())


@dataclass
class AlignBatchResult:
    """parasail Result fields for n problems in structure-of-arrays form."""
    score: np.ndarray
    end_query: np.ndarray
    end_ref: np.ndarray
    beg_query: np.ndarray
    beg_ref: np.ndarray
    cigar_off: np.ndarray  # int64[n+1]
    cigar_rle: np.ndarray  # uint32[total], len << 4 | op (BAM op codes)

    def runs(self, p: int) -> np.ndarray:
        return self.cigar_rle[self.cigar_off[p]: self.cigar_off[p + 1]]

    def cigar_string(self, p: int) -> str:
        """'12=1X3I' -- the text parasail's `cigar.decode` carries."""
        r = self.runs(p)
        return "".join(f"{int(x) >> 4}{OP_CHARS[int(x) & 15]}" for x in r)

    def expanded(self, p: int) -> str:
        """one letter per alignment column (the reference's `my_cigar`, my_classes.py:428-429)."""
        r = self.runs(p)
        if r.size == 0:
            return ""
        return np.repeat(_OP_LUT[r & 15], r >> 4).tobytes().decode("ascii")


def _ptr(a: np.ndarray | None):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


class Engine:
    """One engine = one GPU (one `smx_handle`).  Not thread-safe; use one per host thread."""

    def __init__(self, device: int = 0, stream: int | None = None):
        self._lib = _lib.load()
        h = ctypes.c_void_p()
        rc = self._lib.smx_create(ctypes.byref(h), int(device), ctypes.c_void_p(stream) if stream else None)
        if rc != 0:
            raise EngineError(f"smx_create failed ({rc}): {self._lib.smx_last_error(None).decode()}")
        self._h = h
        self.device = device
        self._pool_len = 0
        self._siblings = []

    def siblings(self, n: int) -> list:
        """`n` more engines on the same device, created on first use and closed with this engine (the batches in
        flight of savemoney.align_all run one handle each)."""
        while len(self._siblings) < n:
            self._siblings.append(Engine(self.device))
        return self._siblings[:n]

    # -- plumbing ---------------------------------------------------------------------------
    def close(self):
        for e in getattr(self, "_siblings", []):
            e.close()
        self._siblings = []
        if getattr(self, "_h", None):
            self._lib.smx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _check(self, rc: int, what: str):
        if rc != 0:
            raise EngineError(f"{what} failed ({rc}): {self._lib.smx_last_error(self._h).decode()}")

    def set_trace_budget(self, nbytes: int):
        self._check(self._lib.smx_set_trace_budget(self._h, int(nbytes)), "smx_set_trace_budget")

    @property
    def guard_errors(self) -> int:
        """SMX_GUARD=1: damaged guard words around the direction-bit regions so far (-1: mode off)"""
        return int(self._lib.smx_debug_guard_errors(self._h))

    @property
    def trace_budget(self) -> int:
        return int(self._lib.smx_get_trace_budget(self._h))

    @property
    def last_run_ms(self) -> float:
        return float(self._lib.smx_last_run_ms(self._h))

    @property
    def last_run_launches(self) -> int:
        return int(self._lib.smx_last_run_launches(self._h))

    # -- sequence pool ----------------------------------------------------------------------
    def upload_pool(self, pool: bytes | np.ndarray):
        arr = np.frombuffer(pool, dtype=np.uint8) if isinstance(pool, (bytes, bytearray)) else np.ascontiguousarray(pool, dtype=np.uint8)
        self._check(self._lib.smx_pool_upload(self._h, _ptr(arr), arr.size), "smx_pool_upload")
        self._pool_len = arr.size

    def upload_sequences(self, seqs: Sequence[str]) -> tuple[np.ndarray, np.ndarray]:
        """Uploads upper-cased sequences back to back; returns (offsets, lengths)."""
        enc = [s.upper().encode("ascii") for s in seqs]
        lens = np.array([len(e) for e in enc], dtype=np.int32)
        offs = np.zeros(len(enc), dtype=np.int64)
        if len(enc) > 1:
            offs[1:] = np.cumsum(lens[:-1], dtype=np.int64)
        self.upload_pool(b"".join(enc))
        return offs, lens

    # -- A8: batched DP + traceback ------------------------------------------------------------
    def align_prepare(self, s1_off, s1_len, s2_off, s2_len, mode, open_: int, extend: int, match: int, mismatch: int):
        s1_off = np.ascontiguousarray(s1_off, dtype=np.int64)
        s2_off = np.ascontiguousarray(s2_off, dtype=np.int64)
        s1_len = np.ascontiguousarray(s1_len, dtype=np.int32)
        s2_len = np.ascontiguousarray(s2_len, dtype=np.int32)
        mode = np.ascontiguousarray(mode, dtype=np.uint8)
        n = s1_off.size
        if not (s1_len.size == s2_off.size == s2_len.size == mode.size == n):
            raise ValueError("descriptor arrays must have equal length")
        self._n = n
        self._check(self._lib.smx_align_prepare(self._h, _ptr(s1_off), _ptr(s1_len), _ptr(s2_off), _ptr(s2_len), _ptr(mode),
                                                n, open_, extend, match, mismatch), "smx_align_prepare")

    def align_run(self) -> int:
        total = ctypes.c_int64(0)
        self._check(self._lib.smx_align_run(self._h, ctypes.byref(total)), "smx_align_run")
        self._total_runs = int(total.value)
        return self._total_runs

    def align_fetch(self) -> AlignBatchResult:
        n = self._n
        out = [np.empty(n, dtype=np.int32) for _ in range(5)]
        off = np.zeros(n + 1, dtype=np.int64)
        rle = np.empty(max(self._total_runs, 1), dtype=np.uint32)
        self._check(self._lib.smx_align_fetch(self._h, *[_ptr(a) for a in out], _ptr(off), _ptr(rle), rle.size), "smx_align_fetch")
        return AlignBatchResult(*out, off, rle[: self._total_runs])

    def align_batch(self, s1_off, s1_len, s2_off, s2_len, mode, open_: int, extend: int, match: int, mismatch: int) -> AlignBatchResult:
        self.align_prepare(s1_off, s1_len, s2_off, s2_len, mode, open_, extend, match, mismatch)
        self.align_run()
        return self.align_fetch()

    def align_kernel_ms(self) -> tuple[np.ndarray, np.ndarray]:
        """(ms[16], cells[14]) of the last align run: forward classes (0..7 int32 warp-per-problem kshift*2+local,
        8/9 int32 CTA-per-problem global/local, 10..13 packed s16x2 with 1/2/4/8 warps per problem), traceback,
        compaction."""
        ms = np.zeros(16, dtype=np.float64)
        cells = np.zeros(14, dtype=np.int64)
        self._check(self._lib.smx_align_kernel_ms(self._h, _ptr(ms), _ptr(cells)), "smx_align_kernel_ms")
        return ms, cells

    @property
    def align_cells(self) -> int:
        return int(self._lib.smx_align_cells(self._h))

    # -- A3: diagonal match-count scan ---------------------------------------------------------
    def scan_prepare(self, ref_off, ref_len, qry_off, qry_len, topology: int) -> np.ndarray:
        ref_off = np.ascontiguousarray(ref_off, dtype=np.int64)
        qry_off = np.ascontiguousarray(qry_off, dtype=np.int64)
        ref_len = np.ascontiguousarray(ref_len, dtype=np.int32)
        qry_len = np.ascontiguousarray(qry_len, dtype=np.int32)
        n = ref_off.size
        counts_off = np.zeros(n + 1, dtype=np.int64)
        self._check(self._lib.smx_scan_prepare(self._h, _ptr(ref_off), _ptr(ref_len), _ptr(qry_off), _ptr(qry_len), n,
                                               int(topology), _ptr(counts_off)), "smx_scan_prepare")
        self._scan_off = counts_off
        return counts_off

    def scan_run(self):
        self._check(self._lib.smx_scan_run(self._h), "smx_scan_run")

    def scan_fetch(self) -> np.ndarray:
        counts = np.empty(max(int(self._scan_off[-1]), 1), dtype=np.int32)
        self._check(self._lib.smx_scan_fetch(self._h, _ptr(counts), counts.size), "smx_scan_fetch")
        return counts[: int(self._scan_off[-1])]

    def scan_batch(self, ref_off, ref_len, qry_off, qry_len, topology: int) -> tuple[np.ndarray, np.ndarray]:
        off = self.scan_prepare(ref_off, ref_len, qry_off, qry_len, topology)
        self.scan_run()
        return self.scan_fetch(), off

    @property
    def scan_cells(self) -> int:
        return int(self._lib.smx_scan_cells(self._h))

    # -- measurement --------------------------------------------------------------------------
    def measure_int_peak(self, iters: int = 2000) -> dict:
        out = (ctypes.c_double * 3)()
        self._check(self._lib.smx_measure_int_peak(self._h, iters, out), "smx_measure_int_peak")
        return {"viaddmax_gops": out[0], "vimax3_gops": out[1], "iadd_lop_gops": out[2]}


class _PipelineResultOps:
    """accessors shared by the one-batch and the segmented result (both provide runs(i), score, new_query_seq_offset)"""
    _OPS = {1: "I", 2: "D", 4: "S", 5: "H", 7: "=", 8: "X"}
    _LUT = np.zeros(16, dtype="S1")
    for _k, _v in _OPS.items():
        _LUT[_k] = _v.encode()

    def cigar_string(self, i: int) -> str:
        """RLE text as MyResult.cigar renders it (ref_query_alignment.py:449-451)."""
        return "".join(f"{int(x) >> 4}{self._OPS[int(x) & 15]}" for x in self.runs(i))

    def my_cigar(self, i: int) -> str:
        r = self.runs(i)
        if r.size == 0:
            return ""
        return np.repeat(self._LUT[r & 15], r >> 4).tobytes().decode("ascii")

    def offset(self, i: int):
        v = int(self.new_query_seq_offset[i])
        return None if v == -2147483648 else v


@dataclass
class PipelineResult(_PipelineResultOps):
    """Output of the alignment stage for P pairs x 2 strands, index = 2 * pair + strand."""
    score: np.ndarray                  # int32[2P]   MyResult.score (0 for empty results)
    new_query_seq_offset: np.ndarray   # int32[2P]   INT32_MIN encodes None
    status: np.ndarray                 # uint8[2P]   0 aligned, 1 empty MyResult(), 2 reference would raise
    cigar_off: np.ndarray              # int64[2P+1]
    cigar_rle: np.ndarray              # uint32[total]
    stats: dict

    def runs(self, i: int) -> np.ndarray:
        return self.cigar_rle[self.cigar_off[i]: self.cigar_off[i + 1]]

    @property
    def nbytes(self) -> int:
        return int(self.score.nbytes + self.new_query_seq_offset.nbytes + self.status.nbytes + self.cigar_off.nbytes + self.cigar_rle.nbytes)


class SegmentedPipelineResult(_PipelineResultOps):
    """The batches of one align_all call, in order.  score / new_query_seq_offset / status / cigar_off are flat arrays
    over all pair-strands; the CIGAR runs stay in their per-batch arrays (runs(i) slices the right one) and are only
    concatenated if somebody asks for the flat `cigar_rle` (hundreds of MB for 20 000 reads)."""

    def __init__(self, parts):
        self.parts = list(parts)
        self.score = np.concatenate([p.score for p in self.parts])
        self.new_query_seq_offset = np.concatenate([p.new_query_seq_offset for p in self.parts])
        self.status = np.concatenate([p.status for p in self.parts])
        self._first = np.cumsum([0] + [len(p.score) for p in self.parts])          # first pair-strand of every batch
        self._run_base = np.cumsum([0] + [int(p.cigar_off[-1]) for p in self.parts])  # first run of every batch
        self.cigar_off = np.concatenate([self.parts[0].cigar_off] + [p.cigar_off[1:] + b for p, b in zip(self.parts[1:], self._run_base[1:])])
        self.stats = {k: sum(p.stats[k] for p in self.parts) for k in self.parts[0].stats}
        self._flat = None

    def runs(self, i: int) -> np.ndarray:
        k = int(np.searchsorted(self._first, i, side="right")) - 1
        p = self.parts[k]
        j = i - int(self._first[k])
        return p.cigar_rle[p.cigar_off[j]: p.cigar_off[j + 1]]

    @property
    def cigar_rle(self) -> np.ndarray:
        if self._flat is None:
            self._flat = np.concatenate([p.cigar_rle for p in self.parts])
        return self._flat

    @property
    def nbytes(self) -> int:
        return int(self.score.nbytes + self.new_query_seq_offset.nbytes + self.status.nbytes + self.cigar_off.nbytes
                   + sum(p.cigar_rle.nbytes for p in self.parts))


class ShardedPipelineResult(_PipelineResultOps):
    """One alignment stage computed on several GPUs: `units` (reads in post_analysis, plasmid pairs in pre_survey) were
    dealt to the shards, every shard holds `per_unit` consecutive pair-strands per unit; this view restores the
    caller's order (my_fastq.keys(), post_analysis_core.py:63-65) without moving the CIGAR runs.

        parts[k]         result of shard k (PipelineResult or SegmentedPipelineResult)
        unit_index[k]    the units of shard k, ascending (shard.deal)
    """

    def __init__(self, parts, unit_index, per_unit: int):
        self.parts = list(parts)
        self.per_unit = int(per_unit)
        n_units = sum(len(ix) for ix in unit_index)
        self._shard = np.empty(n_units, dtype=np.int32)
        self._local = np.empty(n_units, dtype=np.int64)
        for k, ix in enumerate(unit_index):
            ix = np.asarray(ix, dtype=np.int64)
            self._shard[ix] = k
            self._local[ix] = np.arange(len(ix))
        n = n_units * self.per_unit
        self.score = np.empty(n, np.int32)
        self.new_query_seq_offset = np.empty(n, np.int32)
        self.status = np.empty(n, np.uint8)
        lens = np.empty(n, np.int64)
        for k, (p, ix) in enumerate(zip(self.parts, unit_index)):
            if len(ix) == 0:
                continue
            dst = (np.asarray(ix, dtype=np.int64)[:, None] * self.per_unit + np.arange(self.per_unit)[None, :]).ravel()
            self.score[dst] = p.score
            self.new_query_seq_offset[dst] = p.new_query_seq_offset
            self.status[dst] = p.status
            lens[dst] = np.diff(p.cigar_off)
        self.cigar_off = np.zeros(n + 1, np.int64)
        np.cumsum(lens, out=self.cigar_off[1:])
        self.stats = {}
        for p in self.parts:
            for key, v in p.stats.items():
                self.stats[key] = self.stats.get(key, 0) + v
        self._flat = None

    def runs(self, i: int) -> np.ndarray:
        u, j = divmod(int(i), self.per_unit)
        return self.parts[int(self._shard[u])].runs(int(self._local[u]) * self.per_unit + j)

    @property
    def cigar_rle(self) -> np.ndarray:
        """all runs in the caller's order (one block copy per unit: the pair-strands of a unit are consecutive on both sides)"""
        if self._flat is None:
            flat = np.empty(int(self.cigar_off[-1]), np.uint32)
            per = self.per_unit
            for k, p in enumerate(self.parts):
                src, soff = p.cigar_rle, p.cigar_off
                units = np.nonzero(self._shard == k)[0]
                for u in units:
                    j = int(self._local[u])
                    a, b = int(soff[j * per]), int(soff[(j + 1) * per])
                    g = int(self.cigar_off[u * per])
                    flat[g:g + (b - a)] = src[a:b]
            self._flat = flat
        return self._flat

    @property
    def nbytes(self) -> int:
        return int(sum(p.nbytes for p in self.parts))


_STAT_KEYS = ("t_total", "t_pool", "t_scan", "t_select", "t_chain", "t_dp", "t_stitch", "gpu_scan_ms", "gpu_select_ms",
              "gpu_dp_ms", "scan_cells", "dp_cells", "n_dp_problems", "n_host_redecisions", "n_failed", "kernel_launches",
              "n_pair_strands") + tuple(f"dp_ms_class{c}" for c in range(14)) + ("dp_ms_traceback", "dp_ms_compact") \
    + tuple(f"dp_cells_class{c}" for c in range(14)) + tuple(f"dp_launches_{c}" for c in range(16)) \
    + ("gpu_select_only_ms", "gpu_chain_ms", "dp_cells_ck", "n_dp_problems_ck")


def pack_slice(packed, lo, hi):
    """the sub-batch [lo, hi) of sequences packed by `_pack`, without copying the bases"""
    buf, offs, lens = packed
    base = int(offs[lo])
    end = int(offs[hi - 1]) + int(lens[hi - 1])
    return buf[base:end], offs[lo:hi] - base, lens[lo:hi]


def _pack(seqs):
    """sequences back to back as bytes + offsets + lengths (one join and one encode for a list of str)"""
    if all(isinstance(s, str) for s in seqs):
        lens = np.fromiter(map(len, seqs), dtype=np.int32, count=len(seqs))
        buf = "".join(seqs).encode("ascii")
    else:
        enc = [s.encode("ascii") if isinstance(s, str) else bytes(s) for s in seqs]
        lens = np.array([len(e) for e in enc], dtype=np.int32)
        buf = b"".join(enc)
    offs = np.zeros(len(lens), dtype=np.int64)
    if len(lens) > 1:
        offs[1:] = np.cumsum(lens[:-1], dtype=np.int64)
    return np.frombuffer(buf, dtype=np.uint8), offs, lens


def pipeline_run(self, refs, queries, sigmas, open_: int, extend: int, match: int, mismatch: int, topology: int,
                 omit_too_long: bool, pairs=None, n_threads: int = 0) -> PipelineResult:
    """A1: scan -> select -> chain -> DP -> stitch for every (query, reference) pair and both strands.
    refs/queries: upper-case str (or bytes); sigmas: float64 per reference; pairs: optional int array [P, 2]
    of (ref_idx, query_idx); default = all combinations, query-major."""
    rb, ro, rl = refs if isinstance(refs, tuple) else _pack(refs)
    qb, qo, ql = queries if isinstance(queries, tuple) else _pack(queries)   # tuples: already packed (bytes, offsets, lengths)
    n_refs, n_queries = len(rl), len(ql)
    sig = np.ascontiguousarray(sigmas, dtype=np.float64)
    if sig.size != n_refs:
        raise ValueError("one sigma per reference")
    if pairs is not None:
        pairs = np.asarray(pairs, dtype=np.int32).reshape(-1, 2)
        pr = np.ascontiguousarray(pairs[:, 0]); pq = np.ascontiguousarray(pairs[:, 1]); n_pairs = pairs.shape[0]
    else:
        pr = pq = None; n_pairs = n_refs * n_queries
    total = ctypes.c_int64(0)
    self._check(self._lib.smx_pipeline_run(self._h, _ptr(rb), _ptr(ro), _ptr(rl), n_refs, _ptr(sig), _ptr(qb), _ptr(qo), _ptr(ql),
                                           n_queries, _ptr(pr), _ptr(pq), n_pairs, open_, extend, match, mismatch, int(topology),
                                           int(bool(omit_too_long)), int(n_threads), ctypes.byref(total)), "smx_pipeline_run")
    n = 2 * n_pairs
    score = np.empty(n, np.int32); off = np.empty(n, np.int32); status = np.empty(n, np.uint8)
    coff = np.zeros(n + 1, np.int64); rle = np.empty(max(int(total.value), 1), np.uint32)
    self._check(self._lib.smx_pipeline_fetch(self._h, _ptr(score), _ptr(off), _ptr(status), _ptr(coff), _ptr(rle), rle.size), "smx_pipeline_fetch")
    st = (ctypes.c_double * len(_STAT_KEYS))()
    self._check(self._lib.smx_pipeline_stats(self._h, st, len(_STAT_KEYS)), "smx_pipeline_stats")
    return PipelineResult(score, off, status, coff, rle[: int(total.value)], dict(zip(_STAT_KEYS, list(st))))


Engine.pipeline_run = pipeline_run
