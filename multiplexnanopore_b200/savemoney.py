"""Host-side mirror of the reference interface for the alignment path (same names, argument meaning
and error behaviour), backed by the CUDA engine:

  execute_alignment_core(ref_seq_list, my_fastq, param_dict)   post_analysis/post_analysis_core.py:56-65
  set_distance_matrix(ref_seq_list, param_dict)                 pre_survey/pre_survey_core.py:237-245
  calc_distance(ref_seq_1, ref_seq_2, param_dict)               pre_survey/pre_survey_core.py:251-271
  MyResult                                                      modules/ref_query_alignment.py:425-529

A maintainer switches the reference over by replacing the bodies of those two driver functions with
calls into this module; kwargs, result_dict layout and output files stay unchanged.
"""
from __future__ import annotations

import os
import re
from itertools import combinations

import numpy as np
from scipy import stats

from .engine import Engine

PERCENTILE_FACTOR = 0.1  # MyOptimizedAligner.percentile_factor (ref_query_alignment.py:189)
_RUNS = re.compile(r"((.)\2*)")


class MyResult:
    """Same public fields as the reference's MyResult: my_cigar, score, new_query_seq_offset."""
    __slots__ = ("_src", "_idx", "_my_cigar", "score", "new_query_seq_offset")

    def __init__(self, src=None, idx=0):
        self._src, self._idx, self._my_cigar = src, idx, None
        if src is None:
            self._my_cigar, self.score, self.new_query_seq_offset = "", 0, None
        else:
            self.score = int(src.score[idx])
            self.new_query_seq_offset = src.offset(idx)

    @property
    def my_cigar(self) -> str:
        if self._my_cigar is None:
            self._my_cigar = self._src.my_cigar(self._idx)
        return self._my_cigar

    @my_cigar.setter
    def my_cigar(self, v):
        self._my_cigar = v

    @property
    def cigar(self) -> str:
        if self._src is not None and self._my_cigar is None:
            return self._src.cigar_string(self._idx)
        return "".join(f"{len(run)}{L}" for run, L in _RUNS.findall(self.my_cigar))

    def __len__(self):
        return len(self.my_cigar)

    def to_dict(self):
        return {"cigar": self.cigar, "score": self.score, "new_query_seq_offset": self.new_query_seq_offset}

    def calc_levenshtein_distance(self):
        return len(self.my_cigar) - self.my_cigar.count("=")


def _seq_of(x) -> str:
    """the bases of a str / MySeq / MyRefSeq.  Case is normalised where the sequences are consumed: the pipeline's pool
    kernel upper-cases on the device (MySeq upper-cases on construction, my_classes.py:271-272), `Engine.upload_sequences`
    on upload -- not by a host pass over every read here."""
    return x if isinstance(x, str) else (getattr(x, "_seq", None) or str(x))


def _check_params(p):
    for k in ("gap_open_penalty", "gap_extend_penalty", "match_score", "mismatch_score", "topology_of_dna"):
        if k not in p:
            raise Exception(f"missing key in `param_dict`: {k}")
    if p["topology_of_dna"] not in (0, 1):
        raise Exception(f"unknown topology_of_dna: {p['topology_of_dna']}")


def sigma_for(n_ref: int) -> float:
    return float(stats.norm.ppf(1 - 1 / n_ref * PERCENTILE_FACTOR, loc=0, scale=1))  # :200-201


_default_engines = {}


def default_engine(device: int = 0) -> Engine:
    """the process-wide engine of `device` (created on first use)"""
    if device not in _default_engines:
        _default_engines[device] = Engine(device)
    return _default_engines[device]


def _merge(parts):
    from .engine import SegmentedPipelineResult
    return parts[0] if len(parts) == 1 else SegmentedPipelineResult(parts)


def _engines_for(engine, n):
    """`n` engines on the device of `engine`: the caller's engine plus lazily created siblings, owned by that engine
    (`engine.close()` closes them).  Several handles in flight let the host stages (planning, stitching) of one batch
    overlap the GPU stages of the others, and let kernels of different batches fill each other's tails (measured best
    on B200: 6 to 8 x 1000 reads; 8 tolerates a host with few cores per GPU better)."""
    pool = engine.siblings(n - 1)
    pool = [engine] + pool
    if n > 6:
        # every handle grows its direction-bit arena up to its budget (16 GiB by default): more than six handles would
        # not fit 180 GB next to the other workspaces, so the handles share 100 GiB (smaller chunks per batch)
        for e in pool[:n]:
            e.set_trace_budget(max(1 << 30, (100 << 30) // n))
    return pool[:n]


def visible_devices() -> list:
    from .engine import device_count
    return list(range(device_count()))


def _align_sharded(refs, queries, param_dict, omit_too_long, pairs, devices, n_threads, batch_queries, overlap):
    """Row (e): the units of the stage (reads; plasmid pairs when `pairs` is given) are length-binned by predicted work
    and dealt to the GPUs (shard.deal), every GPU runs its share on its own set of engine handles from its own host
    thread (ctypes releases the GIL), and the results come back in the caller's order.  No collective: the only
    inter-GPU step is this host-side gather.  Replaces multiprocessing.Pool(n_cpu).imap
    (post_analysis_core.py:58-61, pre_survey_core.py:240-244)."""
    from concurrent.futures import ThreadPoolExecutor
    from . import shard
    from .engine import ShardedPipelineResult
    refs = [_seq_of(r) for r in refs]
    queries = [_seq_of(q) for q in queries]
    if pairs is None:
        total_ref = sum(len(r) for r in refs)
        weights = [len(q) * total_ref for q in queries]
        per_unit = 2 * len(refs)
    else:
        pairs = np.asarray(pairs, dtype=np.int32).reshape(-1, 2)
        weights = [len(refs[i]) * len(queries[j]) for i, j in pairs]
        per_unit = 2
    index = shard.deal(weights, len(devices))
    if n_threads <= 0:
        cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 4)
        # one call per GPU in pair mode; in read mode the handles in flight of a GPU are not all in their host stages at once
        n_threads = max(1, min(8, cores // len(devices))) if pairs is not None else max(1, min(4, (cores // len(devices)) // 2))

    def work(k):
        if len(index[k]) == 0:
            return None
        eng = default_engine(devices[k])
        if pairs is None:
            return align_all(refs, [queries[i] for i in index[k]], param_dict, omit_too_long, engine=eng,
                             n_threads=n_threads, batch_queries=batch_queries, overlap=overlap)
        return align_all(refs, queries, param_dict, omit_too_long, pairs=pairs[index[k]], engine=eng, n_threads=n_threads)

    with ThreadPoolExecutor(max_workers=len(devices)) as ex:
        parts = list(ex.map(work, range(len(devices))))
    keep = [k for k, p in enumerate(parts) if p is not None]
    return ShardedPipelineResult([parts[k] for k in keep], [index[k] for k in keep], per_unit)


def align_all(refs, queries, param_dict, omit_too_long, pairs=None, engine=None, n_threads=0, batch_queries=1000,
              overlap=8, devices=None):
    """The alignment stage for every (query, reference) pair and both strands (or the explicit `pairs`).
    Queries are processed in batches of `batch_queries` so that device workspaces stay bounded; `overlap`
    batches are in flight at once (one engine handle and one Python thread each; ctypes drops the GIL).
    `engine` pins the call to one handle's GPU; otherwise `devices` (default: every visible GPU) share the work."""
    if engine is None:
        devs = visible_devices() if devices is None else list(devices)
        if len(devs) > 1 and (len(queries) if pairs is None else len(pairs)) >= 2 * len(devs):
            _check_params(param_dict)
            return _align_sharded(refs, queries, param_dict, omit_too_long, pairs, devs, n_threads, batch_queries, overlap)
        engine = default_engine(devs[0] if devs else 0)
    _check_params(param_dict)
    eng = engine or default_engine()
    refs = [_seq_of(r) for r in refs]
    queries = [_seq_of(q) for q in queries]
    sig = [sigma_for(len(r)) for r in refs]
    args = (param_dict["gap_open_penalty"], param_dict["gap_extend_penalty"], param_dict["match_score"],
            param_dict["mismatch_score"], param_dict["topology_of_dna"], omit_too_long)
    if pairs is None and overlap > 1 and batch_queries < len(queries) < batch_queries * overlap and os.environ.get("SMX_ADAPT_BATCH", "1") != "0":
        # few reads for this GPU (one GPU's share of a strong-scaled call): smaller batches keep more handles busy
        batch_queries = min(batch_queries, max(int(os.environ.get("SMX_ADAPT_MIN", "500")), -(-len(queries) // overlap)))
    if pairs is not None or len(queries) <= batch_queries:
        return eng.pipeline_run(refs, queries, sig, *args, pairs=pairs, n_threads=n_threads)
    from .engine import _pack
    refs = _pack(refs)
    if overlap <= 1:
        starts = list(range(0, len(queries), batch_queries))
        parts = [eng.pipeline_run(refs, _pack(queries[b:b + batch_queries]), sig, *args, n_threads=n_threads) for b in starts]
        return _merge(parts)
    bounds = _batch_bounds(len(queries), batch_queries, overlap)
    import queue
    import threading
    from concurrent.futures import ThreadPoolExecutor
    engines = _engines_for(eng, min(overlap, len(bounds) - 1))
    if n_threads <= 0:  # host stages need about three cores per GPU; every handle in flight runs its own pool
        cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 4)
        n_threads = max(1, min(4, cores // len(engines)))
    free = queue.SimpleQueue()
    for e in engines:
        free.put(e)
    # the batches are packed (str -> bytes + offsets) by one producer thread while the first ones are already on the GPU
    n_batches = len(bounds) - 1
    packed = [None] * n_batches
    ready = [threading.Event() for _ in range(n_batches)]

    def producer():
        for k in range(n_batches):
            try:
                packed[k] = _pack(queries[bounds[k]:bounds[k + 1]])
            except Exception as exc:  # surfaces in the worker that waits for this batch
                packed[k] = exc
            ready[k].set()

    threading.Thread(target=producer, daemon=True).start()

    def work(k):
        ready[k].wait()
        if isinstance(packed[k], Exception):
            raise packed[k]
        e = free.get()
        try:
            return e.pipeline_run(refs, packed[k], sig, *args, n_threads=n_threads)
        finally:
            packed[k] = None
            free.put(e)

    with ThreadPoolExecutor(max_workers=len(engines)) as ex:
        parts = list(ex.map(work, range(n_batches)))
    return _merge(parts)


def _batch_bounds(n, batch, overlap):
    """Batch boundaries for `n` queries.  With several batches in flight the first and the last batches are smaller
    (a quarter, then half a batch): the first forward-DP launch starts sooner and the last handles drain together
    instead of one full batch after the other."""
    head = [batch // 4] * min(overlap // 2, 3) + [batch // 2] * min(overlap // 2, 3)
    tail = head[::-1]
    if n < sum(head) + sum(tail) + batch or min(head, default=0) < 1 or os.environ.get("SMX_RAMP", "1") == "0":
        return list(range(0, n, batch)) + [n]
    bounds, pos = [0], 0
    for sz in head:
        pos += sz; bounds.append(pos)
    end_tail = n - sum(tail)
    while pos + batch <= end_tail:
        pos += batch; bounds.append(pos)
    if pos < end_tail:
        pos = end_tail; bounds.append(pos)
    for sz in tail:
        pos += sz; bounds.append(pos)
    return bounds


def _raise_failed(res):
    bad = np.where(res.status == 2)[0]
    if bad.size:
        raise AssertionError(f"alignment failed an internal assertion of the reference algorithm on pair-strand(s) {bad[:8].tolist()}")


def execute_alignment_core(ref_seq_list, my_fastq, param_dict, engine=None, devices=None):
    """result_dict[query_id] = [MyResult] * (2 * N_ref), order [ref0, ref0_rc, ref1, ...] (:63-65, :85).
    Reads are sharded over `devices` (default: every visible GPU; `engine` pins the call to one handle)."""
    ids = list(my_fastq.keys())
    if not ids:
        _check_params(param_dict)
        return {}  # an empty FASTQ maps to an empty result_dict (the reference's pool.map over no reads)
    reads = [my_fastq[i][0] for i in ids]
    res = align_all(ref_seq_list, reads, param_dict, omit_too_long=True, engine=engine, n_threads=0, devices=devices)
    _raise_failed(res)
    n2 = 2 * len(ref_seq_list)
    out = {}
    for qi, qid in enumerate(ids):
        row = []
        for j in range(n2):
            i = qi * n2 + j
            row.append(MyResult(res, i) if res.status[i] == 0 else MyResult())
        out[qid] = row
    return out


def calc_distance_from(res, p, len_i, len_j) -> int:
    s0, s1 = int(res.score[2 * p]), int(res.score[2 * p + 1])
    if s0 <= 0 and s1 <= 0:
        return len_i + len_j
    i = 2 * p if s0 >= s1 else 2 * p + 1
    r = res.runs(i)
    total = int((r >> 4).sum())
    eq = int((r >> 4)[(r & 15) == 7].sum())
    return total - eq


def calc_distance(ref_seq_1, ref_seq_2, param_dict, engine=None) -> int:
    res = align_all([ref_seq_1], [ref_seq_2], param_dict, omit_too_long=False, engine=engine)
    _raise_failed(res)
    return calc_distance_from(res, 0, len(_seq_of(ref_seq_1)), len(_seq_of(ref_seq_2)))


def set_distance_matrix(ref_seq_list, param_dict, engine=None, devices=None) -> np.ndarray:
    """all i < j pairs: reference = plasmid i, query = plasmid j and its reverse complement; the pairs are sharded
    over `devices` (default: every visible GPU)."""
    seqs = [_seq_of(r) for r in ref_seq_list]
    n = len(seqs)
    pairs = np.array(list(combinations(range(n), 2)), dtype=np.int32).reshape(-1, 2)
    dm = np.zeros((n, n), dtype=int)
    if len(pairs) == 0:
        return dm
    res = align_all(seqs, seqs, param_dict, omit_too_long=False, pairs=pairs, engine=engine, devices=devices)
    _raise_failed(res)
    for p, (i, j) in enumerate(pairs):
        dm[i, j] = dm[j, i] = calc_distance_from(res, p, len(seqs[i]), len(seqs[j]))
    assert (dm >= 0).all()
    return dm


from .outputs import QueryAssignment, IntermediateResults, save_intermediate_results  # noqa: E402,F401  (row N2)


class DenseResult:
    """Result of one legacy dense alignment (the fields the legacy MyResult reads off parasail's Result:
    Modules/1_alignment_consensus_core.py:297-330): score, cigar (RLE text, rendered on first use), beg/end on query
    and reference."""
    __slots__ = ("score", "_runs", "_cigar", "beg_query", "beg_ref", "end_query", "end_ref")

    def __init__(self, score, runs, beg_query, beg_ref, end_query, end_ref):
        self.score, self._runs, self._cigar = score, runs, None
        self.beg_query, self.beg_ref, self.end_query, self.end_ref = beg_query, beg_ref, end_query, end_ref

    @property
    def cigar(self) -> str:
        if self._cigar is None:
            from .engine import OP_CHARS
            self._cigar = "".join(f"{int(x) >> 4}{OP_CHARS[int(x) & 15]}" for x in self._runs)
        return self._cigar


def legacy_dense_align(ref_seq_list, my_fastq, param_dict, engine=None, batch_cells=int(7e10), stats=None, overlap=2,
                       trace_budget=40 << 30):
    """Row N4, the pre-0.2 alignment mode kept in the reference under Modules/ (MyAligner.align_all /
    MyAlignerLinear.align_all, Modules/1_alignment_consensus_core.py:219-296): one local alignment
    `parasail.sw_trace(read, reference + reference)` (circular; `reference` alone for linear topology) per read,
    plasmid and strand, no seeding.  Returns {query_id: [DenseResult(ref0), DenseResult(ref0, rc), DenseResult(ref1), ...]}.
    Runs through `Engine.align_batch` (mode 3, the packed local kernel), `batch_cells` DP cells per call, `overlap` calls in
    flight on handles of their own (the latency-bound traceback and the host side of one call under the forward phase of
    the next).  Dense problems are large (100 M cells, 56 MB of direction bits each): the handles get `trace_budget` bytes
    of arena for the call so that a launch holds enough problems to fill the GPU with narrow teams.  `stats` (a dict) receives the cells, launches and CUDA-event kernel times (bench.py --config legacy_dense)."""
    import threading
    from concurrent.futures import ThreadPoolExecutor
    import queue
    from .engine import MODE_SW
    _check_params(param_dict)
    eng = engine or default_engine()
    circular = param_dict["topology_of_dna"] == 0
    refs = [_seq_of(r).upper() for r in ref_seq_list]
    targets = [r + r if circular else r for r in refs]
    ids = list(my_fastq.keys())
    comp = str.maketrans("ACGTMRWSYKVHDBN", "TGCAKYWSRMBDHVN")
    go, ge = param_dict["gap_open_penalty"], param_dict["gap_extend_penalty"]
    ma, mi = param_dict["match_score"], param_dict["mismatch_score"]
    if stats is not None:
        stats.update(cells=0, launches=0, problems=0, out_bytes=0, kernel_ms_total=0.0, class_ms=[0.0] * 16, class_cells=[0] * 14,
                     class_launches=[0] * 16)
    lock = threading.Lock()
    nt, target_len = len(targets), sum(len(t) for t in targets)
    ranges, pos = [], 0
    while pos < len(ids):   # reads of a call: as many as fit the cell budget (at least one)
        end, cells = pos, 0
        while end < len(ids):
            c = 2 * len(_seq_of(my_fastq[ids[end]][0])) * target_len
            if end > pos and cells + c > batch_cells:
                break
            cells += c
            end += 1
        ranges.append((pos, end))
        pos = end
    engines = [eng] + (eng.siblings(min(overlap, len(ranges)) - 1) if overlap > 1 and len(ranges) > 1 else [])
    free = queue.SimpleQueue()
    budgets = [e.trace_budget for e in engines]
    for e in engines:
        e.set_trace_budget(max(e.trace_budget, int(trace_budget)))
        free.put(e)

    def one_call(rng_):
        lo, hi = rng_
        e = free.get()
        try:
            seqs = list(targets)
            for qid in ids[lo:hi]:
                q = _seq_of(my_fastq[qid][0]).upper()
                seqs += [q, q.translate(comp)[::-1]]
            offs, lens = e.upload_sequences(seqs)
            k = np.arange(hi - lo)
            s1 = (nt + 2 * k[:, None, None] + np.arange(2)[None, None, :] + np.zeros((1, nt, 1), np.int64)).ravel()
            s2 = (np.zeros((hi - lo, 1, 1), np.int64) + np.arange(nt)[None, :, None] + np.zeros((1, 1, 2), np.int64)).ravel()
            res = e.align_batch(offs[s1], lens[s1], offs[s2], lens[s2], np.full(len(s1), MODE_SW, np.uint8), go, ge, ma, mi)
            if stats is not None:
                kms, kcells = e.align_kernel_ms()
                with lock:
                    stats["cells"] += e.align_cells; stats["launches"] += e.last_run_launches; stats["problems"] += len(s1)
                    stats["out_bytes"] += int(res.cigar_rle.nbytes + 5 * res.score.nbytes + res.cigar_off.nbytes)
                    stats["kernel_ms_total"] += float(kms.sum())
                    for c in range(16):
                        stats["class_ms"][c] += float(kms[c]); stats["class_launches"][c] += int(kms[c] > 0)
                    for c in range(14):
                        stats["class_cells"][c] += int(kcells[c])
        finally:
            free.put(e)
        part, p = {}, 0
        for qid in ids[lo:hi]:
            row = []
            for _ in range(2 * nt):
                row.append(DenseResult(int(res.score[p]), res.runs(p), int(res.beg_query[p]), int(res.beg_ref[p]),
                                       int(res.end_query[p]), int(res.end_ref[p])))
                p += 1
            part[qid] = row
        return part

    out = {}
    try:
        if len(engines) == 1:
            parts = [one_call(r) for r in ranges]
        else:
            with ThreadPoolExecutor(max_workers=len(engines)) as ex:
                parts = list(ex.map(one_call, ranges))
    finally:
        for e, b in zip(engines, budgets):
            e.set_trace_budget(b)
    for part in parts:
        out.update(part)
    return out
