"""T1: the CPU oracle obeys the reference's runtime invariants and is optimal."""
import numpy as np
import pytest

from oracle import oracle as orc
from tests import util


def affine_score(ops, go, ge, ma, mi, q, r):
    """score of an expanded CIGAR with parasail's matrix (wildcards score 0)."""
    s, qi, ri, prev = 0, 0, 0, ""
    for c in ops:
        if c in "=X":
            a, b = q[qi], r[ri]
            if a in "ACGT" and b in "ACGT":
                s += ma if a == b else mi
            assert (a == b) == (c == "=")
            qi += 1; ri += 1
        elif c == "I":
            s -= ge if prev == "I" else go
            qi += 1
        elif c == "D":
            s -= ge if prev == "D" else go
            ri += 1
        prev = c
    return s, qi, ri


@pytest.mark.parametrize("scoring", [(3, 1, 1, -2), (5, 2, 2, -3), (1, 1, 1, -1)])
def test_oracle_paths_are_valid_and_optimal(scoring):
    go, ge, ma, mi = scoring
    rng = np.random.default_rng(5)
    probs = util.make_problems(rng, 150, 90, alphabet=b"ACGTN")
    for q, r in probs:
        for mode in range(4):
            o = orc.align(mode, q, r, go, ge, ma, mi)
            assert o.score == orc.score_only(mode, q, r, go, ge, ma, mi)
            if mode == orc.MODE_NW:
                s, qi, ri = affine_score(o.ops, go, ge, ma, mi, q, r)
                assert (s, qi, ri) == (o.score, len(q), len(r))
            elif mode == orc.MODE_SW:
                qs, rs = q[o.beg_query: o.end_query + 1], r[o.beg_ref: o.end_ref + 1]
                s, qi, ri = affine_score(o.ops, go, ge, ma, mi, qs, rs)
                if o.score > 0:
                    assert (s, qi, ri) == (o.score, len(qs), len(rs))
            else:
                # semi-global CIGARs cover both sequences completely (asserts at
                # ref_query_alignment.py:70-71,84-85 and the re-clipping at :562-622 rely on it)
                _, qi, ri = affine_score(o.ops, go, ge, ma, mi, q, r)
                assert (qi, ri) == (len(q), len(r))
                assert (o.beg_query, o.beg_ref) == (0, 0)
                if mode == orc.MODE_SG_QB_DB:
                    assert (o.end_query, o.end_ref) == (len(q) - 1, len(r) - 1)


def test_sg_free_ends_score():
    # query is a prefix of the reference: sg_qe_de must not pay for the unaligned reference tail
    o = orc.align(orc.MODE_SG_QE_DE, "ACGTACGT", "ACGTACGTTTTTTTTT", 3, 1, 1, -2)
    assert o.score == 8 and o.cigar_rle == "8=8D" and (o.end_query, o.end_ref) == (7, 7)
    o = orc.align(orc.MODE_SG_QB_DB, "ACGTACGT", "TTTTTTTTACGTACGT", 3, 1, 1, -2)
    assert o.score == 8 and o.cigar_rle == "8D8="


def test_tie_rules_are_the_documented_ones():
    # E == F > H_dag: Appendix A.3 says DEL (F, printed 'I') wins over INS (E, printed 'D')
    o = orc.align(orc.MODE_NW, "AC", "CA", 1, 1, 1, -5)
    assert o.score == orc.score_only(orc.MODE_NW, "AC", "CA", 1, 1, 1, -5)
    assert o.cigar_rle in ("1I1=1D", "1D1=1I")
    # matrix: wildcard row/column are zero
    m = orc.parasail_shim().matrix_create("ACGT", 1, -2).matrix
    assert m.shape == (5, 5) and (m[4] == 0).all() and (m[:, 4] == 0).all() and m[0, 0] == 1 and m[0, 1] == -2


def test_scan_port_matches_numpy():
    rng = np.random.default_rng(6)
    ref = np.array([ord(c) for c in util.rand_seq(rng, 300)], dtype=np.int64)
    q = np.array([ord(c) for c in util.rand_seq(rng, 450)], dtype=np.int64)
    rep = np.tile(ref, 3)
    want = np.array([(rep[k:k + q.size] == q).sum() for k in range(ref.size)])
    assert np.array_equal(orc.kmer_offset_scan(rep, q, ref.size, q.size), want)
