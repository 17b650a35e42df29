"""Row N2: the consumers of the alignment stage that the reference feeds with one Python
object per pair-strand -- read assignment and the two files written right after the alignment -- fed from the
engine's arrays instead.

  QueryAssignment          msa.QueryAssignment (savemoney/modules/msa.py:28-92): the normalised score table and the
                           assignment come from `smx_assign` (one CUDA thread per read); `save_scores` writes
                           <stem>.summary_scores.csv in the reference's layout without pandas
  IntermediateResults      post_analysis_core.IntermediateResults.save (post_analysis/post_analysis_core.py:89-147): the text
                           blocks of <stem>.intermediate_results.txt are formatted by `smx_ir_format` (host threads in
                           libsmx.so) from the run-length CIGAR arrays and DEFLATEd in parallel into the `.ir` zip the
                           reference's `IntermediateResults.load` reads back
Same names and argument meaning as the reference; outputs are byte-identical to the reference's for identical
alignment results (tests/test_outputs_*.py against files written by the reference's own classes).
"""
from __future__ import annotations

import ctypes
import struct
import time
import zlib
from concurrent.futures import ThreadPoolExecutor
from datetime import datetime
from pathlib import Path

import numpy as np

from . import _lib

HEADER = "SAVEMONEY: ver_0.3.7\nwritten by MU"   # MyHeader (my_classes.py:155-158) with the reference's __about__ strings


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _name_of(ref_seq, i):
    p = getattr(ref_seq, "path", None)
    return p.name if p is not None else f"ref{i}"


class QueryAssignment:
    """Score normalisation and read assignment on the device.

        qa = QueryAssignment(ref_seq_list, my_fastq_or_query_ids, pipeline_result, engine=None)
        qa.set_assignment(score_threshold)      # -> assignment_table [N_query, 3]
        qa.save_scores(save_dir)                # <combined_name_stem>.summary_scores.csv

    Public fields as in the reference: score_table, normalized_score_table (masked where a pair-strand has no result),
    normalized_score_table_summary, assignment_table (classified_ref_seq_idx, is_reverse_compliment, assigned),
    score_threshold, N_query, N_ref_seq_list."""
    query_assignment_version = "qa_0.2.0"

    def __init__(self, ref_seq_list, my_fastq, res, engine=None, ref_names=None, combined_name_stem=None):
        from .savemoney import _seq_of
        n2 = 2 * len(ref_seq_list)
        self._setup([len(_seq_of(r)) for r in ref_seq_list], list(my_fastq.keys()) if hasattr(my_fastq, "keys") else list(my_fastq),
                    np.ascontiguousarray(res.score, dtype=np.int32).reshape(-1, n2), (np.asarray(res.status) != 0).reshape(-1, n2), engine)
        self.ref_names = list(ref_names) if ref_names is not None else [_name_of(r, i) for i, r in enumerate(ref_seq_list)]
        self.combined_name_stem = combined_name_stem or getattr(my_fastq, "combined_name_stem", "reads")

    @classmethod
    def from_tables(cls, ref_lens, query_ids, score_table, empty_mask, engine=None, ref_names=None, combined_name_stem="reads"):
        self = cls.__new__(cls)
        self._setup(list(ref_lens), list(query_ids), np.ascontiguousarray(score_table, dtype=np.int32), np.asarray(empty_mask, dtype=bool), engine)
        self.ref_names = list(ref_names) if ref_names is not None else [f"ref{i}" for i in range(len(self.ref_lens))]
        self.combined_name_stem = combined_name_stem
        return self

    def _setup(self, ref_lens, query_ids, score, empty, engine):
        if score.shape != (len(query_ids), 2 * len(ref_lens)) or empty.shape != score.shape:
            raise ValueError("score table must be [N_query, 2 * N_ref] (forward, reverse complement per reference)")
        self.ref_lens, self.query_ids, self._engine = ref_lens, query_ids, engine
        self.score_table = score.astype(int)
        self._score32 = np.ascontiguousarray(score, dtype=np.int32)
        self._status = np.ascontiguousarray(empty, dtype=np.uint8)
        self.score_threshold = None
        self.assignment_table = np.empty((len(query_ids), 3), dtype=int)
        self._run(0.0)   # the tables that do not depend on the threshold

    def _run(self, threshold):
        from .savemoney import default_engine
        eng = self._engine or default_engine()
        n, r = len(self.query_ids), len(self.ref_lens)
        asg = np.empty((n, 3), dtype=np.int32)
        norm = np.empty((n, 2 * r), dtype=np.float64)
        summ = np.empty((n, r), dtype=np.float64)
        lens = np.ascontiguousarray(self.ref_lens, dtype=np.int32)
        eng._check(eng._lib.smx_assign(eng._h, _ptr(self._score32), _ptr(self._status), _ptr(lens), n, r, float(threshold),
                                       _ptr(asg), _ptr(norm), _ptr(summ)), "smx_assign")
        self.normalized_score_table = np.ma.masked_array(norm, mask=self._status.astype(bool))
        self.normalized_score_table_summary = np.ma.masked_array(summ, mask=False)
        return asg

    @property
    def N_query(self):
        return len(self.query_ids)

    @property
    def N_ref_seq_list(self):
        return len(self.ref_lens)

    def set_assignment(self, score_threshold):
        self.score_threshold = score_threshold
        self.assignment_table[:] = self._run(score_threshold)
        return self.assignment_table

    def assigned_groups(self):
        """per reference: [(query_index, is_rc)] of the reads assigned to it, in read order (what iter_assignment_info,
        msa.py:65-78, selects)"""
        out = [[] for _ in self.ref_lens]
        ok = np.nonzero(self.assignment_table[:, 2] == 1)[0]
        for qi in ok:
            out[int(self.assignment_table[qi, 0])].append((int(qi), int(self.assignment_table[qi, 1])))
        return out

    def save_scores(self, save_dir, ref_names=None, combined_name_stem=None):
        """<stem>.summary_scores.csv (msa.py:79-92): tab-separated, one row per read."""
        names = list(ref_names) if ref_names is not None else self.ref_names
        stem = combined_name_stem or self.combined_name_stem
        head = ["query_idx", "query_id"]
        head += [f"{n} (idx={i}{t})" for i, n in enumerate(names) for t in ("", ",rc")]
        head += [f"{n} (idx={i}{t}, normalized)" for i, n in enumerate(names) for t in ("", ",rc")]
        head += ["classified_ref_seq_idx", "is_reverse_compliment", "assigned"]
        norm = np.ma.getdata(self.normalized_score_table)
        lines = ["\t".join(head)]
        for qi, qid in enumerate(self.query_ids):
            row = [str(qi), _csv_field(str(qid))]
            row += [str(int(v)) for v in self.score_table[qi]]
            row += ["NaN" if v != v else repr(float(v)) for v in norm[qi]]
            row += [str(int(v)) for v in self.assignment_table[qi]]
            lines.append("\t".join(row))
        path = Path(save_dir) / f"{stem}.summary_scores.csv"
        path.write_text("\n".join(lines) + "\n")
        return path


def _csv_field(s: str) -> str:
    # csv.QUOTE_MINIMAL as pandas.to_csv applies it with sep="\t"
    if any(c in s for c in '\t"\n\r'):
        return '"' + s.replace('"', '""') + '"'
    return s


# ------------------------------------------------------------------------------------------------------------------
# <stem>.intermediate_results.ir
# ------------------------------------------------------------------------------------------------------------------
def format_result_blocks(res, first_index=0, n_threads=4) -> bytes:
    """the `# resultK(dict)` blocks of one (Segmented / Sharded)PipelineResult, K counted from `first_index`"""
    lib = _lib.load()
    parts = getattr(res, "parts", None)
    if parts is not None and hasattr(res, "_shard"):      # sharded over GPUs: the blocks follow the caller's read order
        return _format_arrays(lib, res.score, res.new_query_seq_offset, res.status, res.cigar_off, res.cigar_rle, first_index, n_threads)
    if parts is not None:                                  # batches of one GPU, already in order
        out, k = [], first_index
        for p in parts:
            out.append(_format_arrays(lib, p.score, p.new_query_seq_offset, p.status, p.cigar_off, p.cigar_rle, k, n_threads))
            k += len(p.score)
        return b"".join(out)
    return _format_arrays(lib, res.score, res.new_query_seq_offset, res.status, res.cigar_off, res.cigar_rle, first_index, n_threads)


def _format_arrays(lib, score, offset, status, cigar_off, cigar_rle, first_index, n_threads) -> bytes:
    score = np.ascontiguousarray(score, dtype=np.int32)
    offset = np.ascontiguousarray(offset, dtype=np.int32)
    status = np.ascontiguousarray(status, dtype=np.uint8)
    cigar_off = np.ascontiguousarray(cigar_off, dtype=np.int64)
    cigar_rle = np.ascontiguousarray(cigar_rle, dtype=np.uint32)
    n = len(score)
    if cigar_rle.size == 0:
        cigar_rle = np.zeros(1, np.uint32)
    need = lib.smx_ir_format(_ptr(score), _ptr(offset), _ptr(status), _ptr(cigar_off), _ptr(cigar_rle), n, first_index, n_threads, None, 0)
    if need < 0:
        raise ValueError("smx_ir_format rejected its arguments")
    buf = np.empty(max(int(need), 1), dtype=np.uint8)
    got = lib.smx_ir_format(_ptr(score), _ptr(offset), _ptr(status), _ptr(cigar_off), _ptr(cigar_rle), n, first_index, n_threads, _ptr(buf), int(need))
    assert got == need
    return buf[: int(need)].tobytes()


def _block(key, kind, payload: str) -> bytes:
    return f"# {key}({kind})\n{payload}\n\n".encode("utf-8")


class IntermediateResults:
    """Writer for the checkpoint the reference loads with IntermediateResults.load (post_analysis_core.py:137-145).

        ir = IntermediateResults(ref_seq_list, my_fastq, param_dict, pipeline_result)
        ir.save(save_dir / f"{my_fastq.combined_name_stem}.intermediate_results.ir")

    ref_seq_list / my_fastq are the reference's own objects (`.path.name`, `.my_hash`, `.path`, `.keys()`), or plain
    stand-ins given through the keyword arguments."""
    intermediate_results_version = "ir_0.2.0"

    def __init__(self, ref_seq_list, my_fastq, param_dict, res, *, ref_seq_names=None, ref_seq_hash_list=None, my_fastq_names=None,
                 my_fastq_hash=None, query_id_list=None, when=None):
        self.header = HEADER + f"\nintermediate_results_version: {self.intermediate_results_version}"
        self.datetime = when or datetime.now()
        self.ref_seq_names = list(ref_seq_names) if ref_seq_names is not None else [r.path.name for r in ref_seq_list]
        self.ref_seq_hash_list = list(ref_seq_hash_list) if ref_seq_hash_list is not None else [r.my_hash for r in ref_seq_list]
        self.my_fastq_names = list(my_fastq_names) if my_fastq_names is not None else [p.name for p in my_fastq.path]
        self.my_fastq_hash = my_fastq_hash if my_fastq_hash is not None else my_fastq.my_hash
        self.param_dict = dict(param_dict)
        self.query_id_list = list(query_id_list) if query_id_list is not None else list(my_fastq.keys())
        self.res = res
        n = len(self.query_id_list) * len(self.ref_seq_names) * 2
        if len(res.score) != n:   # the reference's assertion at post_analysis_core.py:128
            raise AssertionError(f"{len(res.score)} results for {len(self.query_id_list)} reads x {len(self.ref_seq_names)} references x 2 strands")

    def to_bytes(self, n_threads=4) -> bytes:
        head = b"".join([
            _block("header", "str", self.header), _block("datetime", "str", str(self.datetime)),
            _block("ref_seq_names", "list", "\n".join(self.ref_seq_names)), _block("ref_seq_hash_list", "list", "\n".join(self.ref_seq_hash_list)),
            _block("my_fastq_names", "list", "\n".join(self.my_fastq_names)), _block("my_fastq_hash", "str", str(self.my_fastq_hash)),
            _block("param_dict", "dict", "\n".join(f"{k}\t{v}" for k, v in self.param_dict.items())),
            _block("query_id_list", "list", "\n".join(self.query_id_list))])
        return head + format_result_blocks(self.res, 0, n_threads)

    def to_text(self) -> str:
        return self.to_bytes().decode("utf-8")

    def save(self, save_path, n_threads=8):
        """zip (DEFLATE level 9) with the single member <stem>.intermediate_results.txt (my_classes.py:75-81)"""
        save_path = Path(save_path)
        write_zip_member(save_path, save_path.with_suffix(".txt").name, self.to_bytes(n_threads), n_threads, when=self.datetime)
        return save_path


def save_intermediate_results(ref_seq_list, my_fastq, param_dict, res, save_dir, **kw):
    """execute_alignment's checkpoint step (post_analysis_core.py:51-53) from the engine's result arrays"""
    ir = IntermediateResults(ref_seq_list, my_fastq, param_dict, res, **kw)
    stem = kw.get("combined_name_stem") or getattr(my_fastq, "combined_name_stem", "reads")
    return ir.save(Path(save_dir) / f"{stem}.intermediate_results.ir")


def parallel_deflate(data: bytes, level=9, n_threads=8, chunk=4 << 20) -> bytes:
    """One raw DEFLATE stream made of independently compressed chunks (every chunk but the last ends with a sync flush,
    i.e. on a byte boundary with no final-block bit), so the chunks compress on separate threads (zlib releases the GIL).
    Any inflater reads the concatenation as one stream; the .ir text of 20 000 reads (hundreds of MB) takes seconds
    instead of the half minute a single level-9 stream needs."""
    view = memoryview(data)
    pieces = [view[i:i + chunk] for i in range(0, len(view), chunk)] or [view]

    def one(k):
        c = zlib.compressobj(level, zlib.DEFLATED, -15)
        body = c.compress(pieces[k])
        return body + c.flush(zlib.Z_FINISH if k == len(pieces) - 1 else zlib.Z_SYNC_FLUSH)

    if len(pieces) == 1 or n_threads <= 1:
        return b"".join(one(k) for k in range(len(pieces)))
    with ThreadPoolExecutor(max_workers=n_threads) as ex:
        return b"".join(ex.map(one, range(len(pieces))))


def write_zip_member(path: Path, member: str, data: bytes, n_threads=8, when=None):
    """A one-member zip file (method 8 = DEFLATE; zip64 fields when a size needs them) around parallel_deflate."""
    comp = parallel_deflate(data, 9, n_threads)
    crc = zlib.crc32(data) & 0xffffffff
    t = when or datetime.now()
    dostime = (t.hour << 11) | (t.minute << 5) | (t.second // 2)
    dosdate = ((max(t.year, 1980) - 1980) << 9) | (t.month << 5) | t.day
    name = member.encode("utf-8")
    z64 = len(data) >= 0xffffffff or len(comp) >= 0xffffffff
    extra = struct.pack("<HHQQ", 1, 16, len(data), len(comp)) if z64 else b""
    sz_u, sz_c = (0xffffffff, 0xffffffff) if z64 else (len(data), len(comp))
    ver = 45 if z64 else 20
    flags = 0x800 if any(b > 127 for b in name) else 0
    local = struct.pack("<IHHHHHIIIHH", 0x04034b50, ver, flags, 8, dostime, dosdate, crc, sz_c, sz_u, len(name), len(extra)) + name + extra
    cextra = struct.pack("<HHQQ", 1, 16, len(data), len(comp)) if z64 else b""
    central = struct.pack("<IHHHHHHIIIHHHHHII", 0x02014b50, (3 << 8) | ver, ver, flags, 8, dostime, dosdate, crc, sz_c, sz_u,
                          len(name), len(cextra), 0, 0, 0, 0o600 << 16, 0) + name + cextra
    cd_off = len(local) + len(comp)
    tail = b""
    if z64:
        tail += struct.pack("<IQHHIIQQQQ", 0x06064b50, 44, 45, 45, 0, 0, 1, 1, len(central), cd_off)
        tail += struct.pack("<IIQI", 0x07064b50, 0, cd_off + len(central), 1)
    tail += struct.pack("<IHHHHIIH", 0x06054b50, 0, 0, 1, 1, len(central), 0xffffffff if z64 else cd_off, 0)
    with open(path, "wb") as f:
        f.write(local); f.write(comp); f.write(central); f.write(tail)
