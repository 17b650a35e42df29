"""Row N2 on the device: smx_assign (one CUDA thread per read) == the reference's msa.QueryAssignment on golden score
tables (ties, negatives, empty results; tests/golden/assignment_v1.json from oracle/make_golden_assignment.py), and the
files written from a pipeline run -- summary_scores.csv and the .ir text -- equal the files the reference's own classes
wrote for the same inputs (tests/golden/outputs_v1.json.gz)."""
import gzip
import json
import re
from pathlib import Path

import numpy as np
import pytest

from tests import golden_util

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).parent / "golden"


def test_assignment_kernel_matches_reference_golden(engine):
    from multiplexnanopore_b200.outputs import QueryAssignment
    cases = json.loads((GOLD / "assignment_v1.json").read_text())["cases"]
    assert len(cases) >= 5
    for c in cases:
        score = np.array(c["score"]); empty = np.array(c["empty"], dtype=bool)
        qa = QueryAssignment.from_tables(c["ref_lens"], [f"read{i}" for i in range(score.shape[0])], score, empty, engine=engine)
        assert np.array_equal(np.ma.getdata(qa.normalized_score_table_summary), np.array(c["summary"]))
        assert np.array_equal(np.ma.getmask(qa.normalized_score_table), empty)
        want_norm = score / np.repeat(np.array(c["ref_lens"]), 2)[None, :]
        assert np.array_equal(np.ma.getdata(qa.normalized_score_table)[~empty], want_norm[~empty])   # same float64 bits as numpy's division
        for thr, want in c["thresholds"].items():
            assert np.array_equal(qa.set_assignment(float(thr)), np.array(want)), f"threshold {thr}"


def test_files_from_a_pipeline_run_equal_the_reference_files(engine, tmp_path):
    from multiplexnanopore_b200 import savemoney as sm
    from multiplexnanopore_b200.outputs import IntermediateResults, QueryAssignment
    from tests.test_outputs_cpu import make_ir, strip_datetime
    with gzip.open(GOLD / "outputs_v1.json.gz", "rt") as f:
        g = json.load(f)
    case = golden_util.case(g["case"])
    res = sm.align_all(case["refs"], case["reads"], g["params"], True, engine=engine)
    assert strip_datetime(make_ir(g, res).to_text()) == g["intermediate_results_txt"]
    qa = QueryAssignment(case["refs"], g["query_ids"], res, engine=engine, ref_names=g["fasta_names"], combined_name_stem=g["combined_name_stem"])
    assert qa.set_assignment(g["params"]["score_threshold"]).tolist() == g["assignment_table"]
    path = qa.save_scores(tmp_path)
    assert path.name == f"{g['combined_name_stem']}.summary_scores.csv"
    assert path.read_text() == g["summary_scores_csv"]
    groups = qa.assigned_groups()
    assert sum(len(x) for x in groups) == sum(1 for r in g["assignment_table"] if r[2] == 1)
