"""Generates tests/golden/msa_v1.json.gz: inputs and outputs of the UNMODIFIED reference's MyMSA.generate_msa
(/root/reference/savemoney/modules/msa.py:882-982, through oracle/ref_harness.py).

Two kinds of cases: (1) the reads assigned to each plasmid of golden alignment cases, stacked exactly like
MyMSAligner.execute does (msa.py:617-629); (2) seeded random column CIGARs that consume a random reference (generate_msa
never looks at base equality), built to hit what real alignments rarely hold: leading / trailing I and S blocks, blocks
next to hard clips (the 'O' filler), 'I' before 'S' inside one block (column order), empty read sets, all-H reads.
Run here:   python -m oracle.make_golden_msa
"""
from __future__ import annotations

import gzip
import json
import re
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from oracle import ref_harness as rh  # noqa: E402

OUT = ROOT / "tests" / "golden" / "msa_v1.json.gz"


def expand(cigar_rle: str) -> str:
    return "".join(c * int(n) for n, c in re.findall(r"(\d+)(.)", cigar_rle))


def random_cigar(rng, n_ref, style):
    """a column string with exactly n_ref reference-consuming letters"""
    out, r = [], 0
    if style in ("clipped", "wild") and rng.random() < 0.7:   # H.. S.. head
        h = int(rng.integers(0, n_ref // 3 + 1))
        out.append("H" * h); r += h
        out.append("S" * int(rng.integers(0, 5)))
    if style == "wild" and rng.random() < 0.5:
        out.append("I" * int(rng.integers(1, 4)))
    tail_h = int(rng.integers(0, (n_ref - r) // 3 + 1)) if style in ("clipped", "wild") and rng.random() < 0.7 else 0
    while r < n_ref - tail_h:
        u = rng.random()
        if u < 0.80:
            n = min(int(rng.integers(1, 12)), n_ref - tail_h - r); out.append(("=" if rng.random() < 0.9 else "X") * n); r += n
        elif u < 0.88:
            n = min(int(rng.integers(1, 4)), n_ref - tail_h - r); out.append("D" * n); r += n
        elif u < 0.96:
            out.append("I" * int(rng.integers(1, 5)))
        elif style == "wild":
            out.append(rng.choice(["S", "SI", "IS", "ISI", "SSI", "H"]) if r < n_ref - tail_h - 1 else "S")
            if out[-1] == "H":
                r += 1
        else:
            out.append("I")
    if tail_h or (style == "wild" and rng.random() < 0.5):
        out.append("S" * int(rng.integers(0, 5)))
        if style == "wild" and rng.random() < 0.3:
            out.append("I" * int(rng.integers(1, 3)) + "S")
    out.append("H" * tail_h)
    return "".join(out)


def main():
    ns = rh.load()
    assert ns is not None, "needs /root/reference and oracle/_ref"
    msa, mc = ns.msa, ns.mc
    params = dict(rh.DEFAULT_PARAMS)
    cases = []

    def run(name, ref, qs, scores, cigs):
        m = msa.MyMSA.generate_msa(mc.MySeq(ref), list(qs), [list(s) for s in scores], list(cigs), params)
        cases.append({"name": name, "ref": ref, "queries": list(qs), "q_scores": [list(map(int, s)) for s in scores], "cigars": list(cigs),
                      "ref_seq_aligned": m.ref_seq_aligned, "query_seq_list_aligned": m.query_seq_list_aligned,
                      "q_scores_list_aligned": [list(map(int, s)) for s in m.q_scores_list_aligned], "my_cigar_list_aligned": m.my_cigar_list_aligned})
        print(f"{name}: {len(qs)} reads, {len(ref)} reference bases -> {len(m.ref_seq_aligned)} columns", flush=True)

    g = json.load(gzip.open(ROOT / "tests" / "golden" / "golden_v1.json.gz", "rt"))
    rng = np.random.default_rng(3)
    for case_name in ("c2_small_circular", "c2_small_fragments", "c4_small_linear", "iupac_maps_circular"):
        case = next(c for c in g["cases"] if c["name"] == case_name)
        for p, ref in enumerate(case["refs"]):
            qs, cigs = [], []
            for read, row in zip(case["reads"], case["results"]):
                best = max(range(len(row)), key=lambda k: row[k][1] / len(case["refs"][k // 2]) if row[k][0] else -1e9)
                if best // 2 != p or not row[best][0]:
                    continue
                cig, _, off = row[best]
                seq = read if best % 2 == 0 else str(mc.MySeq(read).reverse_complement())
                qs.append(mc.MySeq.set_offset_core(seq.upper(), off)); cigs.append(expand(cig))
            if qs:
                run(f"{case_name}_plasmid{p}", ref, qs, [[int(x) for x in rng.integers(1, 42, size=len(q))] for q in qs], cigs)
    for k, (style, n_ref, n_reads) in enumerate([("plain", 60, 5), ("clipped", 90, 7), ("wild", 50, 6), ("wild", 120, 9), ("wild", 33, 12),
                                                   ("clipped", 300, 20), ("wild", 200, 40), ("plain", 40, 0), ("wild", 64, 1)]):
        ref = "".join(rng.choice(list("ACGT"), size=n_ref))
        cigs = [random_cigar(rng, n_ref, style) for _ in range(n_reads)]
        if style == "wild" and n_reads > 2:
            cigs[0] = "H" * n_ref                     # a read that covers nothing
            cigs[1] = "S" * 3 + "H" * n_ref           # ... and one that only soft-clips before it
        qs = ["".join(rng.choice(list("ACGT"), size=sum(c.count(x) for x in "=XIS"))) for c in cigs]
        run(f"random_{style}_{k}", ref, qs, [[int(x) for x in rng.integers(1, 42, size=len(q))] for q in qs], cigs)
    with gzip.open(OUT, "wt") as f:
        json.dump({"generator": "oracle/make_golden_msa.py", "cases": cases}, f)
    print("wrote", OUT, OUT.stat().st_size, "bytes,", len(cases), "cases")


if __name__ == "__main__":
    main()
