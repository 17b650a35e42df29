"""The tie rules of oracle/parasail_port.c that no real parasail could pin (parasail is neither installed nor
installable on this pod: profiles/parasail_probe_r02.log) are run-time switches.  Every variant must still be an
optimal, valid alignment; the sg_qe_de end-cell orders must obey their definitions; and the committed sensitivity
report (profiles/tie_sensitivity_r02.json, made by oracle/tie_sensitivity.py with the UNMODIFIED reference Python)
must say what DESIGN.md section 5 claims."""
import json
from pathlib import Path

import numpy as np
import pytest

from oracle import oracle as orc
from tests import util
from tests.test_oracle_cpu import affine_score

ROOT = Path(__file__).resolve().parents[1]


@pytest.fixture(autouse=True)
def _restore_default_rules():
    yield
    orc.set_tie_rules()


@pytest.mark.parametrize("rules", [dict(gap_prefers_extend=0), dict(h_order_diag_f_e=0), dict(sg_end_order=1),
                                   dict(sg_end_order=2), dict(sw_end_order=0)])
def test_every_tie_variant_is_valid_and_optimal(rules):
    rng = np.random.default_rng(11)
    probs = util.make_problems(rng, 120, 70, low_complexity=0.5)
    changed = 0
    for q, r in probs:
        for mode in range(4):
            orc.set_tie_rules()
            base = orc.align(mode, q, r, 3, 1, 1, -2)
            orc.set_tie_rules(**rules)
            o = orc.align(mode, q, r, 3, 1, 1, -2)
            assert o.score == base.score == orc.score_only(mode, q, r, 3, 1, 1, -2)
            if mode == orc.MODE_SW:
                qs, rs = q[o.beg_query: o.end_query + 1], r[o.beg_ref: o.end_ref + 1]
                if o.score > 0:
                    assert affine_score(o.ops, 3, 1, 1, -2, qs, rs) == (o.score, len(qs), len(rs))
            else:
                _, qi, ri = affine_score(o.ops, 3, 1, 1, -2, q, r)
                assert (qi, ri) == (len(q), len(r))
            changed += o.ops != base.ops or (o.end_query, o.end_ref) != (base.end_query, base.end_ref)
    if rules != dict(sg_end_order=1):  # corner-vs-last-row ties are rare: test_sg_end_orders_follow_their_definitions finds them
        assert changed > 0, "the switch must be live (low-complexity inputs hold ties of every kind)"


def test_sg_end_orders_follow_their_definitions():
    # H(last row) and H(last column) hold the same maximum in several cells: query = reference = poly-A
    q, r = "A" * 6, "A" * 6
    orc.set_tie_rules(sg_end_order=0)
    o0 = orc.align(orc.MODE_SG_QE_DE, q, r, 3, 1, 1, -2)
    assert (o0.score, o0.end_query, o0.end_ref) == (6, 5, 5)          # unique maximum: the corner
    # two maxima: corner (5, 5) and last-row cell (5, 2) cannot tie with match > 0; build a tie with a mismatch tail:
    # query ACGTTT vs reference ACGAAA: H(2,2)=3 is the best of the last column?  use the scan definition instead
    rng = np.random.default_rng(3)
    seen = set()
    for _ in range(400):
        q = util.rand_seq(rng, int(rng.integers(2, 9)), b"AC")
        r = util.rand_seq(rng, int(rng.integers(2, 9)), b"AC")
        ends = []
        for order in (0, 1, 2):
            orc.set_tie_rules(sg_end_order=order)
            o = orc.align(orc.MODE_SG_QE_DE, q, r, 3, 1, 1, -2)
            ends.append((o.end_query, o.end_ref))
            assert o.end_query == len(q) - 1 or o.end_ref == len(r) - 1
        m, n = len(q) - 1, len(r) - 1
        if ends[0] != ends[1]:  # only the corner-vs-last-row tie separates orders 0 and 1
            assert ends[1] == (m, n) and ends[0][0] == m and ends[0][1] < n
            seen.add("corner")
        if ends[0] != ends[2]:  # column-first vs row-first
            assert ends[2][0] == m and ends[0][1] == n
            seen.add("rowfirst")
    assert seen == {"corner", "rowfirst"}


def test_sensitivity_report_matches_design_claims():
    rep = json.loads((ROOT / "profiles" / "tie_sensitivity_r02.json").read_text())["variants"]
    assert all(rep["default"][k] == 0 for k in ("cigar", "score", "offset", "assignment"))
    # the end-cell order (the least certain rule) is washed out by the reference's own re-clipping
    for v in ("sg_end_corner_in_column", "sg_end_row_first"):
        assert all(rep[v][k] == 0 for k in ("cigar", "score", "offset", "assignment"))
    # the two in-matrix rules are not cosmetic
    assert rep["gap_open_on_tie"]["cigar"] > 0 and rep["h_order_diag_e_f"]["cigar"] > 0
