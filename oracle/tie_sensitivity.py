"""What is at stake at the parasail boundary while no real parasail is reachable.

TEST INFRASTRUCTURE ONLY.  Runs the UNMODIFIED reference Python (oracle/ref_harness.py) on the inputs of the golden
cases (tests/golden/golden_v1.json.gz) with ONE tie rule of oracle/parasail_port.c flipped at a time and counts how
many reference-level results move against the committed goldens: final CIGARs, scores (the read x plasmid score
matrix), new_query_seq_offset values, and downstream read assignments.  Writes profiles/tie_sensitivity_r02.json.

    python -m oracle.tie_sensitivity            (here; needs /root/reference and oracle/_ref)

The rules (see the header of parasail_port.c):
  gap_open_on_tie        open vs extend tie -> open instead of extend
  h_order_diag_e_f       H source tie between E and F -> E (CIGAR 'D') instead of F (CIGAR 'I')
  sg_end_corner_in_column  sg_qe_de end cell: corner visited with the last column (beats a tied last-row cell)
  sg_end_row_first       sg_qe_de end cell: last row scanned before the last column
"""
from __future__ import annotations

import gzip
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

from oracle import oracle as orc  # noqa: E402
from oracle import ref_harness as rh  # noqa: E402

GOLDEN = ROOT / "tests" / "golden" / "golden_v1.json.gz"
OUT = ROOT / "profiles" / "tie_sensitivity_r02.json"

VARIANTS = {
    "default": dict(),
    "gap_open_on_tie": dict(gap_prefers_extend=0),
    "h_order_diag_e_f": dict(h_order_diag_f_e=0),
    "sg_end_corner_in_column": dict(sg_end_order=1),
    "sg_end_row_first": dict(sg_end_order=2),
}


def assignment(scores, ref_lens, thr=0.3):
    """msa.QueryAssignment.set_assignment (msa.py:52-64) on one read's row: (ref_idx, is_rc, assigned)"""
    norm = [s / ref_lens[k // 2] if c else None for k, (c, s, _) in enumerate(scores)]
    vals = [v for v in norm if v is not None]
    if not vals:
        return (-1, -1, -1)
    best = max(vals)
    idx = [k for k, v in enumerate(norm) if v == best]
    if len(idx) > 1:
        return (-1, -1, 0)
    return (idx[0] // 2, idx[0] % 2, int(best > thr))


def run(ns, case):
    aligners = [ns.rqa.MyOptimizedAligner(ns.mc.MySeq(r), case["params"]) for r in case["refs"]]
    out = []
    for read in case["reads"]:
        res = rh.reference_align_read_with(ns, aligners, read, case["omit_too_long"])
        out.append([[orc.rle(c), int(s), None if o is None else int(o)] for c, s, o in res])
    return out


def main(max_cases=None):
    ns = rh.load()
    assert ns is not None, "needs /root/reference and oracle/_ref (make -C oracle)"
    g = json.load(gzip.open(GOLDEN, "rt"))
    cases = g["cases"][:max_cases] if max_cases else g["cases"]
    report = {"generator": "oracle/tie_sensitivity.py", "golden": GOLDEN.name, "variants": {}}
    for name, kw in VARIANTS.items():
        orc.set_tie_rules(**kw)
        t0 = time.time()
        n = dict(pair_strands=0, nonempty=0, cigar=0, score=0, offset=0, reads=0, assignment=0)
        for case in cases:
            got = run(ns, case)
            ref_lens = [len(r) for r in case["refs"]]
            for grow, wrow in zip(got, case["results"]):
                n["reads"] += 1
                n["assignment"] += assignment(grow, ref_lens) != assignment(wrow, ref_lens)
                for (gc, gs, go_), (wc, ws, wo) in zip(grow, wrow):
                    n["pair_strands"] += 1
                    n["nonempty"] += bool(wc)
                    n["cigar"] += gc != wc
                    n["score"] += gs != ws
                    n["offset"] += go_ != wo
        n["seconds"] = round(time.time() - t0, 1)
        report["variants"][name] = n
        print(name, n, flush=True)
    orc.set_tie_rules()
    assert all(report["variants"]["default"][k] == 0 for k in ("cigar", "score", "offset", "assignment")), \
        "default rules must reproduce the committed goldens"
    OUT.write_text(json.dumps(report, indent=1) + "\n")
    print("wrote", OUT)


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else None)
