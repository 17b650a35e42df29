"""Row N1 on the device: the batched MyAlignerBase operations and the chunk re-alignment of a polish round against
(1) what the UNMODIFIED reference computed (tests/golden/realign_v1.json.gz: 364 polished chunks, 1 300 recorded calls,
2 200 direct calls of the five operations, two scorings) and (2) the CPU oracle on 1e5 chunk-sized problems."""
import gzip
import json
import time
from pathlib import Path

import numpy as np
import pytest

from tests import util

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).parent / "golden" / "realign_v1.json.gz"


def gold():
    with gzip.open(GOLD, "rt") as f:
        return json.load(f)


def params_of(sc):
    return dict(gap_open_penalty=sc[0], gap_extend_penalty=sc[1], match_score=sc[2], mismatch_score=sc[3])


def test_direct_operations_match_the_reference(engine):
    from multiplexnanopore_b200.realign import BatchedAligner
    g = gold()
    by_scoring = {}
    for d in g["direct"]:
        by_scoring.setdefault(tuple(d[5]), []).append(d)
    assert len(by_scoring) == 2
    seen = set()
    for sc, items in by_scoring.items():
        al = BatchedAligner(params_of(sc), engine)
        got = al.run([(d[0], d[1], d[2], d[3]) for d in items])
        for d, c in zip(items, got):
            assert c == d[4], f"operation kind {d[0]}"
            seen.add(d[0])
    assert seen == {0, 1, 2, 3, 4}
    # the per-call names
    d = next(x for x in g["direct"] if x[0] == 3 and tuple(x[5]) == (3, 1, 1, -2))
    assert BatchedAligner(params_of((3, 1, 1, -2)), engine).my_special_sw(d[1], d[2]) == d[4]


def test_polish_rounds_match_the_reference(engine):
    from multiplexnanopore_b200.realign import ChunkRealigner
    g = gold()
    groups = {}
    for rec in g["records"]:   # one batch per (case, plasmid, round state): same alignment length and offsets
        key = (rec["case"], rec["plasmid"], rec["len_ref_seq_aligned"], tuple(rec["query_seq_offset_list_aligned"]), tuple(sorted(rec["params"].items())))
        groups.setdefault(key, []).append(rec)
    n_chunks = 0
    for key, recs in groups.items():
        cr = ChunkRealigner(recs[0]["params"], engine)
        got = cr.realign(recs[0]["len_ref_seq_aligned"], recs[0]["query_seq_offset_list_aligned"],
                         [(r["consensus"], r["q_chunks"], r["c_chunks"], r["s_idx"]) for r in recs])
        for r, row in zip(recs, got):
            assert row == r["my_cigar_list"]
        n_chunks += len(recs)
    assert n_chunks == len(g["records"]) >= 300


def test_hundred_thousand_chunk_problems_vs_oracle(engine):
    """>= 1e5 chunk-sized (<= 260 x 260) problems in ONE batch: NW / sg_qe_de / sg_qb_db rows checked against the CPU
    oracle port on a seeded sample (every 23rd problem), all of them for structural validity; prints problems/s."""
    from multiplexnanopore_b200.realign import BatchedAligner, NW, SG_QE_DE, SG_QB_DB, SPECIAL_DP
    from oracle import savemoney_port as port
    rng = np.random.default_rng(5)
    base = util.make_problems(rng, 2600, 260, related=0.9, low_complexity=0.15)
    ops = []
    for rep in range(40):
        for i, (q, r) in enumerate(base):
            kind = (NW, NW, NW, SG_QE_DE, SG_QB_DB, SPECIAL_DP)[(i + rep) % 6]
            if kind == SPECIAL_DP:
                cut = (7 * i + 13 * rep) % (len(q) + 1)
                ops.append((kind, r, q[cut:], q[:cut]))
            else:
                ops.append((kind, r, q[rep % 3:] or "A"))
    assert len(ops) >= 100000
    p = dict(gap_open_penalty=3, gap_extend_penalty=1, match_score=1, mismatch_score=-2)
    al = BatchedAligner(p, engine)
    al.run_arrays(ops[:5000])   # warm-up (allocations)
    t0 = time.perf_counter()
    off, rle, status = al.run_arrays(ops)
    dt = time.perf_counter() - t0
    print(f"\n{len(ops)} chunk operations in {dt:.3f} s = {len(ops) / dt:,.0f} operations/s ({al.last_cells / 1e9:.2f} G cells, DP kernels {al.last_ms:.1f} ms)")
    assert not status.any()
    from multiplexnanopore_b200.realign import expand_all
    cigs = expand_all(off, rle)
    sc = port.Scoring.of(p)
    for i in range(0, len(ops), 23):
        kind, r, q1 = ops[i][0], ops[i][1], ops[i][2]
        if kind == NW:
            want = port.dp(0, r, q1, sc, None).ops
        elif kind == SG_QE_DE:
            want = port.sg_qe_de(r, q1, sc).my_cigar
        elif kind == SG_QB_DB:
            want = port.sg_qb_db(r, q1, sc).my_cigar
        else:
            want = port.special_dp(q1, ops[i][3], r, sc)
        assert cigs[i] == want, f"operation {i} kind {kind}"
    for c, op in zip(cigs[::7], ops[::7]):   # every result consumes exactly the reference and the query
        nq = len(op[2]) + (len(op[3]) if len(op) > 3 else 0)
        assert sum(c.count(x) for x in "=XDH") == len(op[1]) and sum(c.count(x) for x in "=XIS") == nq
