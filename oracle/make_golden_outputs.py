"""Generates tests/golden/outputs_v1.json.gz: the two files the reference writes right after the alignment stage,
produced by the reference's OWN classes (read-only /root/reference through oracle/ref_harness.py):

  <stem>.intermediate_results.txt   IntermediateResults.to_text (post_analysis/post_analysis_core.py:89-134,
                                     modules/my_classes.py:43-74), the member of the `.ir` zip; the `datetime`
                                     block varies run to run and is removed
  <stem>.summary_scores.csv         msa.QueryAssignment.save_scores (modules/msa.py:79-92)

on a golden alignment case (inputs from tests/golden/golden_v1.json.gz, results recomputed by the reference here and
checked against that fixture).   Run here:   python -m oracle.make_golden_outputs
"""
from __future__ import annotations

import gzip
import importlib
import json
import re
import sys
import tempfile
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from oracle import oracle as orc  # noqa: E402
from oracle import ref_harness as rh  # noqa: E402

OUT = ROOT / "tests" / "golden" / "outputs_v1.json.gz"
CASE = "c2_small_circular"
FULL_PARAMS = dict(rh.DEFAULT_PARAMS)


def strip_datetime(text: str) -> str:
    return re.sub(r"# datetime\(str\)\n.*?\n\n", "", text, count=1, flags=re.S)


def main():
    ns = rh.load()
    assert ns is not None, "needs /root/reference and oracle/_ref"
    pac = importlib.import_module("savemoney.post_analysis.post_analysis_core")
    g = json.load(gzip.open(ROOT / "tests" / "golden" / "golden_v1.json.gz", "rt"))
    case = next(c for c in g["cases"] if c["name"] == CASE)
    rng = np.random.default_rng(8)
    with tempfile.TemporaryDirectory() as td:
        td = Path(td)
        fa_names = [f"plasmid_{i}.fa" for i in range(len(case["refs"]))]
        for n, s in zip(fa_names, case["refs"]):
            (td / n).write_text(f">{n}\n{s}\n")
        ids = [f"@read_{i} runid=golden ch={i % 7}" for i in range(len(case["reads"]))]
        ids[3] = "@read\twith_tab"    # a query id that needs csv quoting
        quals = ["".join(chr(33 + int(q)) for q in rng.integers(5, 41, size=len(r))) for r in case["reads"]]
        fq_text = "".join(f"{i}\n{r}\n+\n{q}\n" for i, r, q in zip(ids, case["reads"], quals))
        (td / "reads_golden.fastq").write_text(fq_text)
        ref_seq_list = [ns.mc.MyRefSeq(td / n) for n in fa_names]
        my_fastq = ns.mc.MyFastQ.combine([ns.mc.MyFastQ(td / "reads_golden.fastq")])
        params = dict(FULL_PARAMS, **case["params"])
        aligners = [ns.rqa.MyOptimizedAligner(r, params) for r in ref_seq_list]
        result_dict = {}
        for (qid, (seq, _)), want in zip(my_fastq.items(), case["results"]):
            q = ns.mc.MySeq(seq)
            q_rc = q.reverse_complement()
            row = []
            for al in aligners:   # execute_alignment_core_loop (post_analysis_core.py:68-87)
                for qq in (q, q_rc):
                    cr = al.calc_conserved_region(qq, omit_too_long=True)
                    row.append(al.execute_alignment_using_conserved_regions(qq, cr) if cr is not None else ns.rqa.MyResult())
            got = [[orc.rle(r.my_cigar), int(r.score), r.new_query_seq_offset] for r in row]
            assert got == want, "the reference no longer reproduces golden_v1"
            result_dict[qid] = row
        ir = pac.IntermediateResults(ref_seq_list, my_fastq, params, result_dict)
        ir_text = strip_datetime(ir.to_text())
        qa = ns.msa.QueryAssignment(ref_seq_list, my_fastq, result_dict)
        qa.set_assignment(params["score_threshold"])
        qa.save_scores(td)
        csv_text = (td / f"{my_fastq.combined_name_stem}.summary_scores.csv").read_text()
        out = {"generator": "oracle/make_golden_outputs.py", "case": CASE, "fasta_names": fa_names, "fastq_name": "reads_golden.fastq",
               "fastq_text": fq_text, "query_ids": list(my_fastq.keys()), "params": params,
               "ref_seq_hash_list": [r.my_hash for r in ref_seq_list], "my_fastq_hash": my_fastq.my_hash,
               "combined_name_stem": my_fastq.combined_name_stem, "intermediate_results_txt": ir_text, "summary_scores_csv": csv_text,
               "assignment_table": qa.assignment_table.tolist()}
    with gzip.open(OUT, "wt") as f:
        json.dump(out, f)
    print("wrote", OUT, OUT.stat().st_size, "bytes;", len(ir_text), "chars of .ir text,", len(csv_text), "of csv")


if __name__ == "__main__":
    main()
