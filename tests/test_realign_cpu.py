"""Row N1 without a GPU: the per-read decisions of exec_chunk_poa as restated in realign.plan_read_chunk produce, for
every chunk the reference polished (tests/golden/realign_v1.json.gz, recorded from the UNMODIFIED reference by
oracle/make_golden_realign.py), exactly the sequence of MyAlignerBase calls the reference made -- same operation, same
arguments, same order -- and the literal all-H / all-D rows."""
import gzip
import json
from pathlib import Path

GOLD = Path(__file__).parent / "golden" / "realign_v1.json.gz"


def gold():
    with gzip.open(GOLD, "rt") as f:
        return json.load(f)


def test_planned_operations_equal_the_reference_calls():
    from multiplexnanopore_b200 import realign
    g = gold()
    assert len(g["records"]) >= 300
    n_ops = 0
    kinds = set()
    for rec in g["records"]:
        planned, literals = [], {}
        for idx, (qa, ca) in enumerate(zip(rec["q_chunks"], rec["c_chunks"])):
            cut = rec["len_ref_seq_aligned"] - rec["query_seq_offset_list_aligned"][idx] - rec["s_idx"]
            if cut < 0:
                cut += rec["len_ref_seq_aligned"]
            p = realign.plan_read_chunk(rec["consensus"], qa, ca, cut)
            if p[0] == "literal":
                literals[idx] = p[1]
            else:
                planned.append([p[0], p[1], p[2], p[3] if len(p) > 3 else ""])
        assert planned == [o[:4] for o in rec["ops"]]
        for idx, lit in literals.items():
            assert rec["my_cigar_list"][idx] == lit
        n_ops += len(planned)
        kinds.update(o[0] for o in rec["ops"])
    assert n_ops >= 1200 and {0, 1, 2, 4} <= kinds


def test_expand_all_slices_columns():
    import numpy as np
    from multiplexnanopore_b200.realign import expand_all
    rle = np.array([(3 << 4) | 7, (1 << 4) | 8, (2 << 4) | 1, (4 << 4) | 5, (1 << 4) | 7], np.uint32)
    off = np.array([0, 2, 2, 4, 5], np.int64)
    assert expand_all(off, rle) == ["===X", "", "IIHHHH", "="]
    assert expand_all(np.array([0, 0], np.int64), np.zeros(0, np.uint32)) == [""]
