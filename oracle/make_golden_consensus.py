"""Generates tests/golden/consensus_v1.json.gz: inputs and outputs of the UNMODIFIED reference's Bayesian consensus --
MyMSA.calculate_consensus (/root/reference/savemoney/modules/msa.py:1749-1812) on multiple alignments built by its own
generate_msa, with its own SequenceBasecallQscorePDF (msa.py:654-733) and the compiled calc_consensus_error_rate
(cython_functions/alignment_functions.pyx:93-132, oracle/_ref), through oracle/ref_harness.py.

The event-count tables are SYNTHETIC (seeded; same file format as the reference's NanoporeStats_*.csv, written to a
temporary file and loaded by the reference's class), so no data of the reference is copied.  Cases: the stacks of
tests/golden/msa_v1.json.gz (real alignment-stage CIGARs), random stacks with IUPAC letters in the reference, and deep
stacks (up to 3 000 reads) whose product chains leave the range of a double and run into the denormal / overflow range of
the x87 format.  Recorded per case: both consensus triples and the p_lists of the columns (float.hex; every 16th column of stacks longer than 600 columns).
Run here:   python -m oracle.make_golden_consensus
"""
from __future__ import annotations

import gzip
import json
import sys
import tempfile
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from oracle import ref_harness as rh  # noqa: E402

OUT = ROOT / "tests" / "golden" / "consensus_v1.json.gz"
BASES = "ATCG-"


def synthetic_table(rng, sharp):
    """event counts per (true base, called base) x q-score -1 .. 41, shaped like basecaller statistics"""
    cols = [f"{B}_{c}" for B in BASES for c in BASES]
    counts = np.zeros((43, 25), dtype=np.int64)
    q = np.arange(1, 42)
    for i, B in enumerate(BASES):
        for j, c in enumerate(BASES):
            k = i * 5 + j
            if c == "-":                       # a deletion (or no insertion) carries no q-score: row -1 only
                counts[0, k] = int(rng.integers(1500, 4000)) if B != "-" else int(rng.integers(800000, 1500000))
                continue
            if B == c:
                w = np.exp(-((q - 22.0) / 9.0) ** 2) * 60000 * sharp
            elif B == "-":
                w = np.exp(-q / 7.0) * 900
            else:
                w = np.exp(-q / (5.0 + 3.0 * rng.random())) * (300 + 400 * rng.random())
            counts[2:, k] = rng.poisson(w)
    return cols, counts


def write_csv(path, cols, counts):
    with open(path, "w") as f:
        f.write("# file_version: 9.9.9\n")
        f.write("\t" + "\t".join(cols) + "\n")
        for r, idx in enumerate(range(-1, 42)):
            f.write(str(idx) + "\t" + "\t".join(str(int(x)) for x in counts[r]) + "\n")


def main():
    ns = rh.load()
    assert ns is not None, "needs /root/reference and oracle/_ref"
    msa, mc = ns.msa, ns.mc
    rng = np.random.default_rng(11)
    tables, pdfs = [], []
    with tempfile.TemporaryDirectory() as td:
        for t, sharp in enumerate((1.0, 0.05)):
            cols, counts = synthetic_table(rng, sharp)
            write_csv(Path(td) / f"t{t}.csv", cols, counts)
            pdfs.append(msa.SequenceBasecallQscorePDF(df_csv_path=Path(td) / f"t{t}.csv"))
            tables.append({"columns": cols, "counts": counts.tolist()})
    cases = []

    def run(name, ref, qs, scores, cigs, table, params):
        msa.MyMSA.sbq_pdf = pdfs[table]
        m = msa.MyMSA.generate_msa(mc.MySeq(ref), list(qs), [list(s) for s in scores], list(cigs), dict(params))
        m.calculate_consensus()
        pw, pwo = m.consensus_params(m.param_dict, make_it_for_CYTHON=True)
        plists = []
        for c, r in enumerate(m.ref_seq_aligned):
            ql, sl = [], []
            for i, cig in enumerate(m.my_cigar_list_aligned):
                if cig[c] not in "HO":
                    ql.append(m.query_seq_list_aligned[i][c]); sl.append(m.q_scores_list_aligned[i][c])
            if ql and len(m.ref_seq_aligned) > 600 and c % 16:
                plists.append(0)          # long stacks: the p_list of every 16th column only (0 = not recorded, None = nobody votes)
            elif ql:
                a = m.sbq_pdf.sbq_pdf_CYTHON.calc_consensus_error_rate("".join(ql).encode("utf-8"), sl, pw[r.upper()])
                b = m.sbq_pdf.sbq_pdf_CYTHON.calc_consensus_error_rate("".join(ql).encode("utf-8"), sl, pwo[r.upper()])
                plists.append([[float(x).hex() for x in a], [float(x).hex() for x in b]])
            else:
                plists.append(None)
        cases.append({"name": name, "table": table, "params": {k: params[k] for k in ("error_rate", "ins_rate")}, "ref": ref, "queries": list(qs),
                      "q_scores": [list(map(int, s)) for s in scores], "cigars": list(cigs),
                      "with_prior": [m.with_prior_consensus_seq, [int(x) for x in m.with_prior_consensus_q_scores], m.with_prior_consensus_my_cigar],
                      "without_prior": [m.without_prior_consensus_seq, [int(x) for x in m.without_prior_consensus_q_scores], m.without_prior_consensus_my_cigar],
                      "p_lists": plists})
        print(f"{name}: {len(qs)} reads, {len(m.ref_seq_aligned)} columns, with prior {m.with_prior_consensus_my_cigar[:40]}...", flush=True)

    base = dict(rh.DEFAULT_PARAMS)
    g = json.load(gzip.open(ROOT / "tests" / "golden" / "msa_v1.json.gz", "rt"))
    for k, case in enumerate(g["cases"]):
        if not case["queries"]:
            continue
        if any(set(q) - set("ACGT") for q in case["queries"]):
            # a read base outside ACGT is a missing key of the reference's C++ maps: an empty vector is indexed and the
            # interpreter dies with a segmentation fault (alignment_functions.pyx:113; seen here on the stacks with N calls)
            print("skipped (the reference crashes on it):", case["name"], flush=True)
            continue
        params = dict(base)
        if k % 3 == 1:
            params.update(error_rate=1e-3, ins_rate=1e-4)
        run("msa_" + case["name"], case["ref"], case["queries"], case["q_scores"], case["cigars"], k % 2, params)

    def noisy_stack(n_ref, n_reads, sub, ins, dele, iupac=False):
        ref = "".join(rng.choice(list("ACGT"), size=n_ref))
        qs, cigs = [], []
        for _ in range(n_reads):
            q, c = [], []
            for b in ref:
                u = rng.random()
                if u < dele:
                    c.append("D")
                elif u < dele + sub:
                    q.append(str(rng.choice([x for x in "ACGT" if x != b]))); c.append("X")
                else:
                    q.append(b); c.append("=")
                if rng.random() < ins:
                    q.append(str(rng.choice(list("ACGT")))); c.append("I")
            qs.append("".join(q)); cigs.append("".join(c))
        if iupac:
            ref = list(ref)
            for pos, letter in zip(rng.choice(n_ref, size=min(8, n_ref), replace=False), "NRYKMSWBDHV"):
                ref[int(pos)] = letter
            ref = "".join(ref)
        scores = [[int(x) for x in rng.integers(1, 42, size=len(q))] for q in qs]
        return ref, qs, scores, cigs

    run("iupac_reference", *noisy_stack(80, 25, 0.05, 0.03, 0.03, iupac=True), 0, base)
    run("deep_400", *noisy_stack(40, 400, 0.04, 0.02, 0.03), 0, base)
    run("deep_1500_flat_table", *noisy_stack(24, 1500, 0.08, 0.03, 0.04), 1, dict(base, error_rate=1e-2, ins_rate=1e-3))
    run("deep_3000", *noisy_stack(12, 3000, 0.03, 0.02, 0.02), 0, base)
    run("half_and_half", *noisy_stack(30, 200, 0.48, 0.0, 0.0), 1, base)     # near-ties between two candidates
    with gzip.open(OUT, "wt") as f:
        json.dump({"generator": "oracle/make_golden_consensus.py", "tables": tables, "cases": cases}, f)
    print("wrote", OUT, OUT.stat().st_size, "bytes,", len(cases), "cases")


if __name__ == "__main__":
    main()
