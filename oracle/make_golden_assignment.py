"""Generates tests/golden/assignment_v1.json by running the UNMODIFIED reference msa.QueryAssignment
(/root/reference/savemoney/modules/msa.py:28-64, through oracle/ref_harness.py) on seeded score tables with
ties, negative scores and empty results.   Run here:  python -m oracle.make_golden_assignment
"""
from __future__ import annotations

import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from oracle import ref_harness as rh  # noqa: E402

OUT = ROOT / "tests" / "golden" / "assignment_v1.json"


class _Res:
    def __init__(self, score, empty):
        self.score, self.my_cigar = int(score), ("" if empty else "=")


class _Fastq(dict):
    pass


def main():
    ns = rh.load()
    if ns is None:
        raise SystemExit("reference not available")
    rng = np.random.default_rng(41)
    cases = []
    for case in range(6):
        n_ref = int(rng.integers(1, 7))
        n_q = int(rng.integers(5, 60))
        ref_lens = [int(x) for x in rng.integers(50, 9000, size=n_ref)]
        if case == 2:
            ref_lens = [ref_lens[0]] * n_ref   # equal lengths: ties in the normalised table are likely
        score = rng.integers(-300, 6000, size=(n_q, 2 * n_ref))
        empty = rng.random((n_q, 2 * n_ref)) < 0.35
        for q in range(n_q):   # force ties, all-empty rows and all-negative rows
            r = rng.random()
            if r < 0.15:
                empty[q, :] = True
            elif r < 0.35 and n_ref > 1:
                a, b = rng.choice(2 * n_ref, size=2, replace=False)
                la, lb = ref_lens[a // 2], ref_lens[b // 2]
                score[q, a], score[q, b] = 3 * la, 3 * lb   # identical normalised maxima
                empty[q, a] = empty[q, b] = False
            elif r < 0.45:
                score[q, :] = -np.abs(score[q, :])
        score[empty] = 0
        refs = ["A" * L for L in ref_lens]
        fq = _Fastq((f"read{q}", ["ACGT", [30] * 4]) for q in range(n_q))
        result_dict = {f"read{q}": [_Res(score[q, j], empty[q, j]) for j in range(2 * n_ref)] for q in range(n_q)}
        qa = ns.msa.QueryAssignment(refs, fq, result_dict)
        out = {"ref_lens": ref_lens, "score": score.tolist(), "empty": empty.astype(int).tolist(), "thresholds": {}}
        out["summary"] = np.ma.getdata(qa.normalized_score_table_summary).tolist()
        for thr in (0.3, 0.8, 2.5):
            qa.set_assignment(thr)
            out["thresholds"][str(thr)] = qa.assignment_table.tolist()
        cases.append(out)
    OUT.write_text(json.dumps({"generator": "oracle/make_golden_assignment.py", "cases": cases}))
    print("wrote", OUT, OUT.stat().st_size, "bytes")


if __name__ == "__main__":
    main()
