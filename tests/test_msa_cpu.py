"""Row N3 without a GPU: the column-parallel restatement of generate_msa that csrc/smx_msa.cuh implements (extra columns
per reference gap from the reads' I / S blocks; every read decides alone given its gap's column types; serial simulation
only for gaps in which an 'I' precedes an 'S') is replayed here in plain Python, step for step like the kernels, and
compared with the outputs of the UNMODIFIED reference (tests/golden/msa_v1.json.gz, oracle/make_golden_msa.py).  It
pins the algorithm; tests/test_msa_gpu.py pins the kernels."""
import gzip
import json
from pathlib import Path

GOLD = Path(__file__).parent / "golden" / "msa_v1.json.gz"


def gold():
    with gzip.open(GOLD, "rt") as f:
        return json.load(f)


def column_parallel_msa(ref, queries, q_scores, cigars):
    R, n = len(ref), len(queries)
    # smx_msa_expand: position of every reference-consuming letter, query index at every gap
    pos, qcnt = [], []
    for cig in cigars:
        p, qc, q, pend_q = [], [], 0, 0
        for idx, L in enumerate(cig):
            if L in "=XDH":
                p.append(idx); qc.append(pend_q)
                if L in "=X":
                    q += 1
                pend_q = q
            else:
                q += 1
        p.append(len(cig)); qc.append(pend_q)
        assert len(p) == R + 1
        pos.append(p); qcnt.append(qc)
    block = lambda i, r: ((0 if r == 0 else pos[i][r - 1] + 1), pos[i][r])
    # smx_msa_gaps (+ the host's serial simulation for gaps with an 'I' before an 'S')
    types = []
    for r in range(R + 1):
        blocks = [cigars[i][slice(*block(i, r))] for i in range(n)]
        if any("IS" in b for b in blocks):
            p, t = [0] * n, ""
            while True:
                cur = [b[k] if k < len(b) else "" for b, k in zip(blocks, p)]
                T = "S" if "S" in cur else "I" if "I" in cur else ""
                if not T:
                    break
                p = [k + (c == T) for k, c in zip(p, cur)]
                t += T
            types.append(t)
        else:
            types.append("S" * max([b.count("S") for b in blocks], default=0) + "I" * max([b.count("I") for b in blocks], default=0))
    # smx_msa_fill: one (read, gap) at a time
    ref_al = "".join("-" * len(types[r]) + (ref[r] if r < R else "") for r in range(R + 1))
    seqs, quals, cigs = [], [], []
    for i in range(n):
        cig, L = cigars[i], len(cigars[i])
        s, qv, c = [], [], []
        for r in range(R + 1):
            bs, be = block(i, r)
            p, q = bs, qcnt[i][r]
            for T in types[r]:
                if p < be and cig[p] == T:
                    c.append(T); s.append(queries[i][q]); qv.append(q_scores[i][q]); p += 1; q += 1
                else:
                    cur = cig[p] if p < L else "Z"
                    prev = cig[p - 1] if p > 0 else (cig[L - 1] if L else "Z")
                    hard = cur == "H" or prev == "H" or (cur == "Z" and L > 0 and cig[0] == "H")
                    c.append("O" if hard else "N"); s.append(" " if hard else "-"); qv.append(-1)
            if r < R:
                q = qcnt[i][r] + (be - bs)
                Lr = cig[be]
                c.append(Lr)
                if Lr in "=X":
                    s.append(queries[i][q]); qv.append(q_scores[i][q])
                else:
                    s.append("-" if Lr == "D" else " "); qv.append(-1)
        seqs.append("".join(s)); quals.append(qv); cigs.append("".join(c))
    return ref_al, seqs, quals, cigs


def test_column_parallel_restatement_equals_reference():
    g = gold()
    assert len(g["cases"]) >= 25
    complex_seen = False
    for c in g["cases"]:
        got = column_parallel_msa(c["ref"], c["queries"], c["q_scores"], c["cigars"])
        assert got[0] == c["ref_seq_aligned"], c["name"]
        assert got[1] == c["query_seq_list_aligned"], c["name"]
        assert got[2] == c["q_scores_list_aligned"], c["name"]
        assert got[3] == c["my_cigar_list_aligned"], c["name"]
        complex_seen |= any("IS" in cig for cig in c["cigars"])
    assert complex_seen, "the goldens must hold a gap whose column order needs the serial simulation"


def test_cigar_runs_round_trip():
    import numpy as np
    from multiplexnanopore_b200.msa import cigar_runs
    r = cigar_runs("===XIIDDSSHH=")
    assert [(int(x) >> 4, int(x) & 15) for x in r] == [(3, 7), (1, 8), (2, 1), (2, 2), (2, 4), (2, 5), (1, 7)]
    assert cigar_runs("").size == 0
    import pytest
    with pytest.raises(Exception):
        cigar_runs("==N=")
