"""The parasail-shaped adapter reads like the reference's own calls (ref_query_alignment.py:35,58,68,82,343) and
returns what the oracle's shim returns; also a polish-sized batch (row N1: ~240 x 240 problems, msa.py:1249-1355)."""
import numpy as np
import pytest

from oracle import oracle as orc
from tests import util

pytestmark = pytest.mark.gpu


def test_per_call_surface(engine):
    from multiplexnanopore_b200 import parasail_compat as parasail
    parasail.use_engine(engine)
    shim = orc.parasail_shim()
    m = parasail.matrix_create("ACGT", 1, -2)
    assert np.array_equal(m.matrix, shim.matrix_create("ACGT", 1, -2).matrix)
    rng = np.random.default_rng(3)
    ref = util.rand_seq(rng, 300)
    q = util.mutate(rng, ref[40:260])
    for name in ("nw_trace", "sg_qe_de_trace", "sg_qb_db_trace", "sw_trace"):
        got = getattr(parasail, name)(q, ref, 3, 1, m)
        want = getattr(shim, name)(q, ref, 3, 1, shim.matrix_create("ACGT", 1, -2))
        assert (got.score, got.end_query, got.end_ref) == (want.score, want.end_query, want.end_ref)
        assert got.cigar.decode == want.cigar.decode
        assert (got.cigar.beg_query, got.cigar.beg_ref) == (want.cigar.beg_query, want.cigar.beg_ref)
    with pytest.raises(ValueError):
        parasail.nw_trace("", ref, 3, 1, m)


def test_polish_sized_batch(engine):
    from multiplexnanopore_b200 import parasail_compat as parasail
    parasail.use_engine(engine)
    rng = np.random.default_rng(4)
    m = parasail.matrix_create("ACGT", 1, -2)
    calls = []
    for _ in range(3000):
        cons = util.rand_seq(rng, 240)
        chunk = util.mutate(rng, cons, 0.03, 0.03, 0.03)
        calls.append((int(rng.integers(0, 4)), chunk, cons))
    got = parasail.align_many(calls, 3, 1, m)
    for (mode, s1, s2), g in list(zip(calls, got))[::17]:
        o = orc.align(mode, s1, s2, 3, 1, 1, -2)
        assert (g.score, g.cigar.decode.decode()) == (o.score, o.cigar_rle)
