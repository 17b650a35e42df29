"""Row A10 on the host (`my_special_DP`, ref_query_alignment.py:89-163): the run-length overlap merge of
smx_stitch.hpp against (1) the column-string implementation in the same header and (2) the CPU oracle port
(oracle/savemoney_port.py: msa_cigars + cumsum_scores), on randomized clipped semi-global results.  No GPU."""
from __future__ import annotations

import ctypes
import re
import subprocess
from pathlib import Path

import numpy as np
import pytest

from oracle import savemoney_port as port

ROOT = Path(__file__).resolve().parent.parent
OPS = {"I": 1, "D": 2, "S": 4, "H": 5, "=": 7, "X": 8}
LET = {v: k for k, v in OPS.items()}


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    so = tmp_path_factory.mktemp("stitch") / "libstitch_harness.so"
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", str(so), str(ROOT / "tests" / "stitch_harness.cpp")])
    lib = ctypes.CDLL(str(so))
    u32p, ll = ctypes.POINTER(ctypes.c_uint32), ctypes.c_longlong
    lib.smx_test_merge.restype = ll
    lib.smx_test_merge.argtypes = [u32p, ll, ll, ll, u32p, ll, ll, ll, ll, ll, ll] + [ctypes.c_int] * 5 + [u32p, ll]
    return lib


def encode(cig: str) -> np.ndarray:
    return np.array([(len(m.group(0)) << 4) | OPS[m.group(0)[0]] for m in re.finditer(r"(.)\1*", cig)], dtype=np.uint32)


def decode(runs) -> str:
    return "".join(LET[int(r) & 15] * (int(r) >> 4) for r in runs)


def merge(lib, kept1, ns1, nh1, kept2, ns2, nh2, n_ref, nq1, nq2, sc, column_strings):
    k1, k2 = encode(kept1), encode(kept2)
    out = np.zeros(len(k1) + len(k2) + 16, np.uint32)
    p = lambda a: a.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32))
    n = lib.smx_test_merge(p(k1), len(k1), ns1, nh1, p(k2), len(k2), ns2, nh2, n_ref, nq1, nq2, sc.go, sc.ge, sc.ma, sc.mi,
                           int(column_strings), p(out), len(out))
    return None if n < 0 else decode(out[:n])


def rand_aligned(rng, n_ref_cols, p_gap, long_runs):
    """random column string over =XID covering exactly n_ref_cols reference columns"""
    out, r = [], 0
    while r < n_ref_cols:
        L = rng.choice(list("=XID"), p=[1 - 3 * p_gap, p_gap, p_gap, p_gap])
        n = int(rng.integers(1, 12 if long_runs else 3))
        if L != "I":
            n = min(n, n_ref_cols - r)
            r += n
        out.append(L * n)
    if rng.random() < 0.3:
        out.append("I" * int(rng.integers(1, 4)))
    return "".join(out)


def oracle_merge(kept1, ns1, nh1, kept2, ns2, nh2, sc):
    """the overlap branch of special_dp (port) on explicit clipped results"""
    c1 = kept1 + "S" * ns1 + "H" * nh1
    c2 = "H" * nh2 + "S" * ns2 + kept2
    n_ref = sum(c1.count(x) for x in "=XDH")
    q = lambda c: "A" * sum(c.count(x) for x in "=XIS")
    a1, a2 = port.msa_cigars("A" * n_ref, [q(c1), q(c2)], [c1, c2])
    left = np.array(port.cumsum_scores(a1, sc))
    right = np.array(port.cumsum_scores(a2[::-1], sc)[::-1])
    cut = int((left + right).argmax())
    n_s1 = len(left) - 1 - cut - sum(a1[cut:].count(c) for c in "DNHO")
    n_s2 = cut - sum(a2[:cut].count(c) for c in "DNHO")
    return (a1[:cut] + "S" * n_s1 + "S" * n_s2 + a2[cut:]).replace("N", "").replace("O", "")


def cases(rng, n):
    for it in range(n):
        n_ref = int(rng.integers(1, 120))
        e1 = int(rng.integers(-1, n_ref))           # end_ref of r1 (-1: nothing aligned)
        b2 = int(rng.integers(0, e1 + 2))           # beg_ref of r2, gap = b2 - e1 - 1 <= 0
        p_gap = float(rng.choice([0.02, 0.1, 0.25]))
        long_runs = bool(rng.integers(0, 2))
        kept1 = rand_aligned(rng, e1 + 1, p_gap, long_runs) if e1 >= 0 else ""
        kept2 = rand_aligned(rng, n_ref - b2, p_gap, long_runs) if b2 < n_ref else ""
        ns1, ns2 = int(rng.integers(0, 6)) * int(rng.integers(0, 2)), int(rng.integers(0, 6)) * int(rng.integers(0, 2))
        yield kept1, ns1, n_ref - e1 - 1, kept2, ns2, b2, n_ref


@pytest.mark.parametrize("scoring", [(3, 1, 1, -2), (5, 2, 2, -3), (1, 1, 1, -1), (4, 0, 2, -6)])
def test_run_length_merge_matches_column_strings_and_oracle(harness, scoring):
    sc = port.Scoring(*scoring)
    rng = np.random.default_rng(sum(scoring) + 77)
    n_checked = 0
    for kept1, ns1, nh1, kept2, ns2, nh2, n_ref in cases(rng, 1500):
        if not kept1 and ns1 == 0 and nh1 == 0:
            continue
        nq1 = sum(kept1.count(x) for x in "=XI") + ns1
        nq2 = sum(kept2.count(x) for x in "=XI") + ns2
        a = merge(harness, kept1, ns1, nh1, kept2, ns2, nh2, n_ref, nq1, nq2, sc, False)
        b = merge(harness, kept1, ns1, nh1, kept2, ns2, nh2, n_ref, nq1, nq2, sc, True)
        assert a == b, (kept1, ns1, nh1, kept2, ns2, nh2)
        if n_checked % 5 == 0 and a is not None:
            assert a == oracle_merge(kept1, ns1, nh1, kept2, ns2, nh2, sc), (kept1, ns1, nh1, kept2, ns2, nh2)
        n_checked += 1
    assert n_checked > 1000


def test_inconsistent_rows_fail_in_both(harness):
    sc = port.Scoring(3, 1, 1, -2)
    # row 2 covers one reference column too few / query count wrong: both implementations refuse
    for kept1, nh1, kept2, nh2, n_ref, dq in [("====", 0, "==", 1, 4, 0), ("====", 0, "===", 1, 4, 1), ("==", 3, "====", 0, 4, 0)]:
        nq1 = sum(kept1.count(x) for x in "=XI")
        nq2 = sum(kept2.count(x) for x in "=XI") + dq
        assert merge(harness, kept1, 0, nh1, kept2, 0, nh2, n_ref, nq1, nq2, sc, False) is None
        assert merge(harness, kept1, 0, nh1, kept2, 0, nh2, n_ref, nq1, nq2, sc, True) is None


def test_long_overlap(harness):
    """reads-sized rows: two nearly full-length alignments of a 9 kb reference overlapping almost completely"""
    sc = port.Scoring(3, 1, 1, -2)
    rng = np.random.default_rng(5)
    for _ in range(20):
        n_ref = int(rng.integers(5000, 9000))
        e1, b2 = n_ref - int(rng.integers(1, 400)), int(rng.integers(0, 400))
        kept1, kept2 = rand_aligned(rng, e1 + 1, 0.03, False), rand_aligned(rng, n_ref - b2, 0.03, False)
        nq1, nq2 = sum(kept1.count(x) for x in "=XI") + 7, sum(kept2.count(x) for x in "=XI")
        a = merge(harness, kept1, 7, n_ref - e1 - 1, kept2, 0, b2, n_ref, nq1, nq2, sc, False)
        b = merge(harness, kept1, 7, n_ref - e1 - 1, kept2, 0, b2, n_ref, nq1, nq2, sc, True)
        assert a is not None and a == b


def test_rotation_on_runs_matches_column_strings(harness):
    """Row A11 (set_ref_seq_offset_as_0, ref_query_alignment.py:465-511): block rotation of the runs against the
    column-string restatement in the same header and the oracle port."""
    ll = ctypes.c_longlong
    harness.smx_test_rotate.restype = ll
    harness.smx_test_rotate.argtypes = [ctypes.POINTER(ctypes.c_uint32), ll, ll, ll, ll, ctypes.c_int, ctypes.POINTER(ll)]
    rng = np.random.default_rng(11)
    n_rot = 0
    for it in range(1500):
        n_ref = int(rng.integers(1, 200)) if it % 25 else int(rng.integers(2000, 6000))  # a few read-sized rows
        cig = rand_aligned(rng, n_ref, float(rng.choice([0.03, 0.2])), bool(rng.integers(0, 2)))
        if rng.random() < 0.5:  # clipped tail / head as the stitcher produces them
            cig = "H" * int(rng.integers(0, 9)) + "S" * int(rng.integers(0, 9)) + cig + "S" * int(rng.integers(0, 9)) + "H" * int(rng.integers(0, 9))
        n_r = sum(cig.count(x) for x in "=XDH")
        n_q = sum(cig.count(x) for x in "=XIS")
        ref_off, q_off = int(rng.integers(0, n_r + 1)) * int(rng.integers(0, 2)), int(rng.integers(0, n_q + 1)) * int(rng.integers(0, 2))
        got = []
        for mode in (0, 1):
            runs = np.zeros(len(cig) + 4, np.uint32)
            enc = encode(cig)
            runs[:len(enc)] = enc
            off = ll(0)
            n = harness.smx_test_rotate(runs.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32)), len(enc), len(runs), ref_off, q_off, mode, ctypes.byref(off))
            assert n >= 0
            got.append((decode(runs[:n]), int(off.value), [int(r) for r in runs[:n]]))
        assert got[0] == got[1], (cig, ref_off, q_off)
        want_cig, want_off = port.rotate_to_ref0(cig, ref_off, q_off)
        assert got[0][0] == want_cig and got[0][1] == want_off, (cig, ref_off, q_off)
        n_rot += want_cig != cig
    assert n_rot > 300


def test_rotation_at_checkpoint_boundaries_and_score(harness):
    """The suffix sums behind the rotation keep one checkpoint per 64 runs: run counts around the multiples of 64 with
    every cut position, and the branch-free score, against the oracle port."""
    ll = ctypes.c_longlong
    harness.smx_test_rotate.restype = ll
    harness.smx_test_rotate.argtypes = [ctypes.POINTER(ctypes.c_uint32), ll, ll, ll, ll, ctypes.c_int, ctypes.POINTER(ll)]
    harness.smx_test_score.restype = ll
    harness.smx_test_score.argtypes = [ctypes.POINTER(ctypes.c_uint32), ll, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int]
    rng = np.random.default_rng(12)
    for n_runs in (1, 2, 63, 64, 65, 127, 128, 129, 191, 192, 193, 300):
        ops = "=X=I=D=X"
        cig = "".join(ops[k % len(ops)] * int(rng.integers(1, 4)) for k in range(n_runs))
        if n_runs > 2:
            cig = "HH" + "S" + cig[3:] if cig[3:] else cig
        enc = encode(cig)
        n_r = sum(cig.count(x) for x in "=XDH"); n_q = sum(cig.count(x) for x in "=XIS")
        for sc in ((3, 1, 1, -2), (5, 2, 2, -3)):
            got = harness.smx_test_score(np.asarray(enc, np.uint32).ctypes.data_as(ctypes.POINTER(ctypes.c_uint32)), len(enc), *sc)
            assert got == port.cigar_score(cig, port.Scoring(go=sc[0], ge=sc[1], ma=sc[2], mi=sc[3]))
        for ref_off in sorted(set([0, 1, n_r // 3, n_r // 2, max(n_r - 1, 0), n_r] + [int(x) for x in rng.integers(0, n_r + 1, size=6)])):
            q_off = int(rng.integers(0, n_q + 1))
            runs = np.zeros(len(enc) + 4, np.uint32)
            runs[:len(enc)] = enc
            off = ll(0)
            n = harness.smx_test_rotate(runs.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32)), len(enc), len(runs), ref_off, q_off, 0, ctypes.byref(off))
            want_cig, want_off = port.rotate_to_ref0(cig, ref_off, q_off)
            assert decode(runs[:n]) == want_cig and int(off.value) == want_off, (n_runs, ref_off, q_off)
