"""Row N2 without a GPU: the `.ir` checkpoint written from result arrays (outputs.IntermediateResults over smx_ir_format,
host threads in libsmx.so) is byte-identical to the text the reference's own IntermediateResults.to_text produced for
the same alignment results (tests/golden/outputs_v1.json.gz, made by oracle/make_golden_outputs.py; the datetime block
is excluded), and the zip container round-trips."""
import gzip
import hashlib
import json
import re
import zipfile
from pathlib import Path

import numpy as np
import pytest

from tests import golden_util

GOLD = Path(__file__).parent / "golden" / "outputs_v1.json.gz"
OPS = {"I": 1, "D": 2, "S": 4, "H": 5, "=": 7, "X": 8}


def gold():
    with gzip.open(GOLD, "rt") as f:
        return json.load(f)


def result_from_golden(case):
    from multiplexnanopore_b200.engine import PipelineResult
    score, off, status, coff, rle = [], [], [], [0], []
    for row in case["results"]:
        for cig, s, o in row:
            runs = [(int(n) << 4) | OPS[c] for n, c in re.findall(r"(\d+)(.)", cig)]
            rle += runs
            coff.append(coff[-1] + len(runs))
            score.append(s)
            off.append(-2147483648 if o is None else o)
            status.append(0 if cig else 1)
    return PipelineResult(np.array(score, np.int32), np.array(off, np.int32), np.array(status, np.uint8), np.array(coff, np.int64),
                          np.array(rle, np.uint32), {})


def strip_datetime(text):
    return re.sub(r"# datetime\(str\)\n.*?\n\n", "", text, count=1, flags=re.S)


def make_ir(g, res):
    from multiplexnanopore_b200.outputs import IntermediateResults
    case = golden_util.case(g["case"])
    # hashes restated: sha256 of the sequence (my_classes.py:353-355) / of the re-serialised FASTQ (:240-247)
    ref_hash = [hashlib.sha256(r.encode()).hexdigest() for r in case["refs"]]
    assert ref_hash == g["ref_seq_hash_list"]
    assert hashlib.sha256(g["fastq_text"].encode()).hexdigest() == g["my_fastq_hash"]   # Q-scores <= 41: to_string() is the file
    return IntermediateResults(None, None, g["params"], res, ref_seq_names=g["fasta_names"], ref_seq_hash_list=ref_hash,
                               my_fastq_names=[g["fastq_name"]], my_fastq_hash=g["my_fastq_hash"], query_id_list=g["query_ids"])


def test_ir_text_equals_reference_text():
    g = gold()
    res = result_from_golden(golden_util.case(g["case"]))
    ir = make_ir(g, res)
    assert strip_datetime(ir.to_text()) == g["intermediate_results_txt"]


def test_ir_blocks_from_batches_and_threads():
    from multiplexnanopore_b200.engine import PipelineResult, SegmentedPipelineResult
    from multiplexnanopore_b200.outputs import format_result_blocks
    g = gold()
    res = result_from_golden(golden_util.case(g["case"]))
    whole = format_result_blocks(res, 0, n_threads=1)
    assert whole == format_result_blocks(res, 0, n_threads=7)
    cut = 60
    a = PipelineResult(res.score[:cut], res.new_query_seq_offset[:cut], res.status[:cut], res.cigar_off[:cut + 1].copy(), res.cigar_rle[:res.cigar_off[cut]], {})
    b = PipelineResult(res.score[cut:], res.new_query_seq_offset[cut:], res.status[cut:], res.cigar_off[cut:] - res.cigar_off[cut], res.cigar_rle[res.cigar_off[cut]:], {})
    assert format_result_blocks(SegmentedPipelineResult([a, b]), 0) == whole
    assert whole.decode().startswith("# result0(dict)\ncigar\t") and f"# result{len(res.score) - 1}(dict)" in whole.decode()


def test_ir_zip_round_trip(tmp_path):
    g = gold()
    res = result_from_golden(golden_util.case(g["case"]))
    ir = make_ir(g, res)
    path = ir.save(tmp_path / f"{g['combined_name_stem']}.intermediate_results.ir", n_threads=3)
    with zipfile.ZipFile(path) as z:
        assert z.testzip() is None
        assert z.namelist() == [f"{g['combined_name_stem']}.intermediate_results.txt"]
        assert z.getinfo(z.namelist()[0]).compress_type == zipfile.ZIP_DEFLATED
        text = z.read(z.namelist()[0]).decode("utf-8")
    assert strip_datetime(text) == g["intermediate_results_txt"]


def test_parallel_deflate_is_one_valid_stream():
    import zlib
    from multiplexnanopore_b200.outputs import parallel_deflate
    rng = np.random.default_rng(0)
    data = (b"12=1X3I" * 1000 + rng.integers(0, 256, size=70000, dtype=np.uint8).tobytes()) * 9
    for chunk in (1 << 12, 1 << 16, 1 << 30):
        comp = parallel_deflate(data, 9, 4, chunk)
        assert zlib.decompress(comp, -15) == data
    assert zlib.decompress(parallel_deflate(b"", 9, 4), -15) == b""


def test_reference_loads_the_checkpoint(tmp_path):
    """when the reference tree is present: its own IntermediateResults.load reads our file back into the same result_dict"""
    from oracle import ref_harness as rh
    if not rh.REF_ROOT.is_dir() or rh.load() is None:
        pytest.skip("/root/reference is not present (GPU box)")
    import importlib
    pac = importlib.import_module("savemoney.post_analysis.post_analysis_core")
    g = gold()
    case = golden_util.case(g["case"])
    res = result_from_golden(case)
    path = make_ir(g, res).save(tmp_path / "reads_golden.intermediate_results.ir")
    ir = pac.IntermediateResults()
    ir.load(path)
    rd = ir.result_dict
    assert list(rd.keys()) == g["query_ids"]
    from oracle import oracle as orc
    for row, want in zip(rd.values(), case["results"]):
        assert [[orc.rle(r.my_cigar), int(r.score), r.new_query_seq_offset] for r in row] == want
    assert ir.param_dict["gap_open_penalty"] == g["params"]["gap_open_penalty"]
