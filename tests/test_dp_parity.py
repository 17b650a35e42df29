"""T0: CUDA DP + traceback == CPU oracle, bit for bit, through the C-ABI.
Covers parasail.{nw,sg_qe_de,sg_qb_db,sw}_trace as called at ref_query_alignment.py:35,58,68,82,343."""
import numpy as np
import pytest

from oracle import oracle as orc
from tests import util

pytestmark = pytest.mark.gpu

SCORINGS = [(3, 1, 1, -2), (5, 2, 2, -3), (2, 2, 1, -1), (4, 1, 3, -1)]  # (open, extend, match, mismatch)


def run_and_compare(engine, probs, modes, scoring):
    go, ge, ma, mi = scoring
    pool, s1o, s1l, s2o, s2l = util.pack_problems(probs)
    engine.upload_pool(pool)
    res = engine.align_batch(s1o, s1l, s2o, s2l, np.asarray(modes, np.uint8), go, ge, ma, mi)
    for p, ((q, r), mode) in enumerate(zip(probs, modes)):
        o = orc.align(int(mode), q, r, go, ge, ma, mi)
        ctx = f"problem {p} mode {mode} m={len(q)} n={len(r)} scoring={scoring}"
        assert res.score[p] == o.score, ctx
        assert (res.end_query[p], res.end_ref[p]) == (o.end_query, o.end_ref), ctx
        assert (res.beg_query[p], res.beg_ref[p]) == (o.beg_query, o.beg_ref), ctx
        assert res.cigar_string(p) == o.cigar_rle, ctx


@pytest.mark.parametrize("scoring", SCORINGS)
def test_small_problems_all_modes(engine, scoring):
    rng = np.random.default_rng(11)
    probs = util.make_problems(rng, 600, 70)
    modes = rng.integers(0, 4, size=len(probs))
    run_and_compare(engine, probs, modes, scoring)


def test_class_boundaries(engine):
    """row counts around the kernel-class and band boundaries (32, 64, 128, 256, 512 rows)."""
    rng = np.random.default_rng(12)
    probs, modes = [], []
    for m in (1, 2, 31, 32, 33, 63, 64, 65, 127, 128, 129, 255, 256, 257, 300, 511, 512, 513, 700):
        for n in (1, 2, 31, 32, 33, 64, 100, 257):
            ref = util.rand_seq(rng, n)
            q = util.mutate(rng, (ref * (m // n + 2))[:m], 0.1, 0.0, 0.0)[:m]
            for mode in range(4):
                probs.append((q, ref)); modes.append(mode)
    run_and_compare(engine, probs, modes, (3, 1, 1, -2))


def test_medium_problems(engine):
    rng = np.random.default_rng(13)
    probs = util.make_problems(rng, 120, 1500, related=0.8)
    modes = rng.integers(0, 3, size=len(probs))
    run_and_compare(engine, probs, modes, (3, 1, 1, -2))


def test_tie_heavy(engine):
    """homopolymers, all-equal scores, go == ge: every tie rule of Appendix A.3/A.6 is exercised."""
    rng = np.random.default_rng(14)
    probs, modes = [], []
    for _ in range(300):
        a = "A" * int(rng.integers(1, 90)) + "C" * int(rng.integers(0, 5)) + "A" * int(rng.integers(0, 60))
        b = "A" * int(rng.integers(1, 90)) + "C" * int(rng.integers(0, 5)) + "A" * int(rng.integers(0, 60))
        probs.append((a, b)); modes.append(int(rng.integers(0, 4)))
    for _ in range(200):
        probs.append((util.rand_seq(rng, int(rng.integers(1, 60)), b"AC"), util.rand_seq(rng, int(rng.integers(1, 60)), b"AC")))
        modes.append(int(rng.integers(0, 4)))
    for scoring in [(1, 1, 1, -1), (2, 2, 2, -2), (3, 1, 1, -2)]:
        run_and_compare(engine, probs, modes, scoring)


def test_wildcards(engine):
    """non-ACGT bytes score 0 against everything but print '=' when equal (Appendix A.5)."""
    rng = np.random.default_rng(15)
    probs = util.make_problems(rng, 200, 120, alphabet=b"ACGTNRY", low_complexity=0.0)
    modes = rng.integers(0, 4, size=len(probs))
    run_and_compare(engine, probs, modes, (3, 1, 1, -2))


def test_large_sg_pair(engine):
    """one problem of the size class that carries 96% of the cells (PROBE: sg_* up to 4.5k x 9.8k)."""
    rng = np.random.default_rng(16)
    ref = util.rand_seq(rng, 6000)
    q = util.mutate(rng, ref[1500:5200])
    probs = [(q, ref), (q, ref), (util.mutate(rng, ref), ref)]
    run_and_compare(engine, probs, [1, 2, 0], (3, 1, 1, -2))


def test_trace_budget_chunking(engine):
    rng = np.random.default_rng(17)
    probs = util.make_problems(rng, 80, 900)
    modes = rng.integers(0, 4, size=len(probs))
    engine.set_trace_budget(1 << 20)
    try:
        run_and_compare(engine, probs, modes, (3, 1, 1, -2))
    finally:
        engine.set_trace_budget(16 << 30)


def test_rejects_bad_input(engine):
    from multiplexnanopore_b200 import EngineError
    engine.upload_pool(b"ACGTACGT")
    with pytest.raises(EngineError):
        engine.align_batch([0], [0], [4], [4], [0], 3, 1, 1, -2)  # empty query
    with pytest.raises(EngineError):
        engine.align_batch([0], [4], [4], [8], [0], 3, 1, 1, -2)  # slice outside the pool
    with pytest.raises(EngineError):
        engine.align_batch([0], [4], [4], [4], [9], 3, 1, 1, -2)  # unknown mode
