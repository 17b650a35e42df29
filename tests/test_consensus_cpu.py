"""Row N3, second half, without a GPU:
(1) the integer restatement of the x87 multiplication that the consensus kernel runs on the device (csrc/smx_ld80.hpp)
    against the FPU's own long double -- normal, denormal and overflowing products, ties, zeros / infinities / NaNs, and
    long product chains;
(2) the oracle's C restatement of calc_consensus_error_rate (oracle/consensus_port.c) plus the host logic of
    multiplexnanopore_b200/consensus.py (statistics table, priors, minimum, mixed bases, q-score, consensus CIGAR) against
    the outputs of the UNMODIFIED reference (tests/golden/consensus_v1.json.gz, oracle/make_golden_consensus.py), on
    columns built by the CPU replay of generate_msa (tests/test_msa_cpu.py)."""
import ctypes
import gzip
import json
import subprocess
from pathlib import Path

import numpy as np
import pytest

from oracle import oracle as orc
from tests.test_msa_cpu import column_parallel_msa

ROOT = Path(__file__).resolve().parent.parent
GOLD = ROOT / "tests" / "golden" / "consensus_v1.json.gz"


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    so = tmp_path_factory.mktemp("cons") / "libconsensus_harness.so"
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", str(so), str(ROOT / "tests" / "consensus_harness.cpp")])
    lib = ctypes.CDLL(str(so))
    lib.smx_test_ldmul.restype = ctypes.c_longlong
    lib.smx_test_ldmul.argtypes = [ctypes.c_ulonglong, ctypes.c_longlong, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
    lib.smx_test_ldchain.restype = ctypes.c_longlong
    lib.smx_test_ldchain.argtypes = [ctypes.c_ulonglong, ctypes.c_longlong, ctypes.c_double]
    return lib


@pytest.mark.parametrize("regime", range(6))
def test_restated_x87_multiplication(harness, regime):
    a, b = np.zeros(1, np.longdouble), np.zeros(1, np.longdouble)
    bad = harness.smx_test_ldmul(100 + regime, 1_000_000, regime, a.ctypes.data, b.ctypes.data)
    assert bad == 0, f"regime {regime}: {bad} products differ from the FPU, first {a[0]!r} * {b[0]!r}"


@pytest.mark.parametrize("centre", [1.0, 1e-3, 1e3, 30.0, 1 / 30.0])
def test_restated_product_chains(harness, centre):
    # 100 000 factors around `centre`: through the whole exponent range into the denormals / to infinity, compared after every step
    assert harness.smx_test_ldchain(7, 100_000, centre) == 0


def gold():
    with gzip.open(GOLD, "rt") as f:
        return json.load(f)


def test_oracle_and_host_logic_equal_the_reference():
    from multiplexnanopore_b200 import consensus as cs
    g = gold()
    pdfs = [cs.SequenceBasecallQscorePDF(columns=t["columns"], counts=t["counts"]) for t in g["tables"]]
    tabs = [p.tables() for p in pdfs]
    assert len(g["cases"]) >= 30
    deep = 0
    for case in g["cases"]:
        ref_al, seqs, quals, cigs = column_parallel_msa(case["ref"], case["queries"], case["q_scores"], case["cigars"])
        params = dict(case["params"])
        with_prior, without_prior = cs.consensus_params(params)
        P, pdf = tabs[case["table"]]
        p_w, p_wo, n_ev = [], [], []
        for c, r in enumerate(ref_al):
            votes = [(seqs[i][c], quals[i][c]) for i in range(len(seqs)) if cigs[i][c] not in "HO"]
            n_ev.append(len(votes))
            rec = case["p_lists"][c]
            if not votes:
                assert rec is None
                p_w.append([0.0] * 5); p_wo.append([0.0] * 5)
                continue
            called = "".join(v[0] for v in votes)
            q = [v[1] for v in votes]
            a = orc.consensus_error_rate(P, pdf, called, q, [with_prior[r.upper()][b] for b in cs.BASES])
            b = orc.consensus_error_rate(P, pdf, called, q, [without_prior[r.upper()][b] for b in cs.BASES])
            if rec:   # recorded columns: bit for bit
                assert [x.hex() for x in a] == rec[0] and [x.hex() for x in b] == rec[1], (case["name"], c)
            p_w.append(a); p_wo.append(b)
        deep = max(deep, max(n_ev))
        got_w = cs._consensus_core(ref_al, p_w, n_ev)
        got_wo = cs._consensus_core(ref_al, p_wo, n_ev)
        assert [got_w[0], [int(x) for x in got_w[1]], got_w[2]] == case["with_prior"], case["name"]
        assert [got_wo[0], [int(x) for x in got_wo[1]], got_wo[2]] == case["without_prior"], case["name"]
    assert deep >= 2500, "the goldens must hold a stack deep enough to leave the range of a double"


def test_oracle_rejects_what_crashes_the_reference():
    from multiplexnanopore_b200 import consensus as cs
    t = gold()["tables"][0]
    P, pdf = cs.SequenceBasecallQscorePDF(columns=t["columns"], counts=t["counts"]).tables()
    with pytest.raises(ValueError):
        orc.consensus_error_rate(P, pdf, "AN", [30, 30], [0.2] * 5)


def _matrices(seqs, quals, cigs, n_cols):
    n = len(seqs)
    seq = np.zeros((n, n_cols), np.uint8); qs = np.zeros((n, n_cols), np.int8); cig = np.zeros((n, n_cols), np.uint8)
    for i in range(n):
        seq[i] = np.frombuffer(seqs[i].encode("ascii"), np.uint8)
        cig[i] = np.frombuffer(cigs[i].encode("ascii"), np.uint8)
        qs[i] = np.asarray(quals[i], np.int8)
    return seq, qs, cig


def test_host_code_of_the_entry_point_around_a_cpu_replay_of_the_kernel(harness):
    """smx_consensus_run's own host code (factor table, start values, final sums: csrc/smx_consensus_host.hpp) and the
    kernel's per-thread body (smx_cons_column, the same function the kernel calls) run on the CPU: p_lists bit for bit and
    both consensus triples of every golden case.  What is left to the GPU test is the launch."""
    from multiplexnanopore_b200 import consensus as cs
    harness.smx_test_consensus.restype = ctypes.c_int
    harness.smx_test_consensus.argtypes = [ctypes.c_void_p] * 3 + [ctypes.c_int, ctypes.c_longlong] + [ctypes.c_void_p] * 5 + [ctypes.c_int] + [ctypes.c_void_p] * 3
    g = gold()
    tabs = [cs.SequenceBasecallQscorePDF(columns=t["columns"], counts=t["counts"]).tables() for t in g["tables"]]
    for case in g["cases"]:
        ref_al, seqs, quals, cigs = column_parallel_msa(case["ref"], case["queries"], case["q_scores"], case["cigars"])
        C = len(ref_al)
        seq, qs, cig = _matrices(seqs, quals, cigs, C)
        with_prior, without_prior = cs.consensus_params(dict(case["params"]))
        letters = sorted(set(ref_al.upper()))
        col_class = np.array([letters.index(ch) for ch in ref_al.upper()], np.int32)
        pn_w = np.array([[with_prior[ch][b] for b in cs.BASES] for ch in letters], np.float64)
        pn_wo = np.array([[without_prior[ch][b] for b in cs.BASES] for ch in letters], np.float64)
        P, pdf = tabs[case["table"]]
        p_w = np.zeros((C, 5)); p_wo = np.zeros((C, 5)); n_ev = np.zeros(C, np.int32)
        rc = harness.smx_test_consensus(seq.ctypes.data, qs.ctypes.data, cig.ctypes.data, len(seqs), C, P.ctypes.data, pdf.ctypes.data, col_class.ctypes.data,
                                        pn_w.ctypes.data, pn_wo.ctypes.data, len(letters), p_w.ctypes.data, p_wo.ctypes.data, n_ev.ctypes.data)
        assert rc == 0, case["name"]
        for c, rec in enumerate(case["p_lists"]):
            if rec is None:
                assert n_ev[c] == 0
            elif rec:
                assert [float(x).hex() for x in p_w[c]] == rec[0] and [float(x).hex() for x in p_wo[c]] == rec[1], (case["name"], c)
        got_w = cs._consensus_core(ref_al, p_w, n_ev)
        got_wo = cs._consensus_core(ref_al, p_wo, n_ev)
        assert [got_w[0], [int(x) for x in got_w[1]], got_w[2]] == case["with_prior"], case["name"]
        assert [got_wo[0], [int(x) for x in got_wo[1]], got_wo[2]] == case["without_prior"], case["name"]
    bad = np.frombuffer(b"AN", np.uint8).reshape(2, 1).copy()
    rc = harness.smx_test_consensus(bad.ctypes.data, np.array([[30], [30]], np.int8).ctypes.data, np.frombuffer(b"==", np.uint8).reshape(2, 1).copy().ctypes.data, 2, 1,
                                    tabs[0][0].ctypes.data, tabs[0][1].ctypes.data, np.zeros(1, np.int32).ctypes.data, np.full((1, 5), 0.2).ctypes.data,
                                    np.full((1, 5), 0.2).ctypes.data, 1, np.zeros((1, 5)).ctypes.data, np.zeros((1, 5)).ctypes.data, np.zeros(1, np.int32).ctypes.data)
    assert rc == -7


def test_statistics_table_file_and_selection(tmp_path):
    """the csv reader (same format as the reference's NanoporeStats_*.csv) gives the tables the constructor-from-counts gives,
    and select_sbq_pdf (MyMSA.set_sbq_pdf, msa.py:1738-1745) picks the table whose q-score distribution is closest"""
    from multiplexnanopore_b200 import consensus as cs
    g = gold()
    paths = []
    for k, t in enumerate(g["tables"]):
        p = tmp_path / f"NanoporeStats_0.{k}.csv"
        with open(p, "w") as f:
            f.write("# file_version: 0.0.%d\n" % k)
            f.write("\t" + "\t".join(t["columns"]) + "\n")
            for idx, row in zip(range(-1, 42), t["counts"]):
                f.write(str(idx) + "\t" + "\t".join(str(x) for x in row) + "\n")
        paths.append(p)
    for p, t in zip(paths, g["tables"]):
        a, b = cs.SequenceBasecallQscorePDF(df_csv_path=p), cs.SequenceBasecallQscorePDF(columns=t["columns"], counts=t["counts"])
        assert a.NanoporeStats_PDF_version.startswith("0.0.") and np.array_equal(a.tables()[0], b.tables()[0]) and np.array_equal(a.tables()[1], b.tables()[1])
        assert abs(sum(a.pdf_core["A_A"]) - 1.0) < 1e-12 and len(a.pdf_core["A_A"]) == 93
    pdfs = [cs.SequenceBasecallQscorePDF(df_csv_path=p) for p in paths]
    for k, pdf in enumerate(pdfs):
        assert cs.select_sbq_pdf(pdf.q_score_distribution, paths).NanoporeStats_PDF_version == f"0.0.{k}"


def test_zero_probabilities_infinities_and_nans_follow_the_fpu(harness):
    """q-scores 0 and 42..92 hit zero entries of pdf_core (msa.py:683-687 smooths the rows 1..41 only; :705 pads with zeros):
    factors 0 / x, x / 0 and 0 / 0.  The CPU replay of the entry point (restated multiplication with its zero / infinity /
    NaN classes) must still equal the oracle's native long double, NaN for NaN."""
    import math
    from multiplexnanopore_b200 import consensus as cs
    harness.smx_test_consensus.restype = ctypes.c_int
    harness.smx_test_consensus.argtypes = [ctypes.c_void_p] * 3 + [ctypes.c_int, ctypes.c_longlong] + [ctypes.c_void_p] * 5 + [ctypes.c_int] + [ctypes.c_void_p] * 3
    t = gold()["tables"][1]
    P, pdf = cs.SequenceBasecallQscorePDF(columns=t["columns"], counts=t["counts"]).tables()
    with_prior, without_prior = cs.consensus_params(dict(error_rate=1e-3, ins_rate=1e-4))
    rng = np.random.default_rng(77)
    R, C = 12, 400
    seq = rng.choice(np.frombuffer(b"ATCG-", np.uint8), size=(R, C), p=[0.23, 0.23, 0.23, 0.23, 0.08])
    qs = rng.choice(np.array([-1, 0, 0, 5, 20, 41, 42, 60, 92], np.int8), size=(R, C))
    qs[seq == ord("-")] = -1
    cig = rng.choice(np.frombuffer(b"==XDHO", np.uint8), size=(R, C))
    ref = rng.choice(np.frombuffer(b"ATCG-N", np.uint8), size=C)
    letters = sorted(set(ref.tobytes().decode()))
    col_class = np.array([letters.index(chr(x)) for x in ref], np.int32)
    pn_w = np.array([[with_prior[ch][b] for b in cs.BASES] for ch in letters], np.float64)
    pn_wo = np.array([[without_prior[ch][b] for b in cs.BASES] for ch in letters], np.float64)
    p_w = np.zeros((C, 5)); p_wo = np.zeros((C, 5)); n_ev = np.zeros(C, np.int32)
    rc = harness.smx_test_consensus(seq.ctypes.data, qs.ctypes.data, cig.ctypes.data, R, C, P.ctypes.data, pdf.ctypes.data, col_class.ctypes.data,
                                    pn_w.ctypes.data, pn_wo.ctypes.data, len(letters), p_w.ctypes.data, p_wo.ctypes.data, n_ev.ctypes.data)
    assert rc == 0
    same = lambda a, b: all((math.isnan(x) and math.isnan(y)) or x == y for x, y in zip(a, b))
    special = 0
    for c in range(C):
        votes = [(chr(seq[r, c]), int(qs[r, c])) for r in range(R) if cig[r, c] not in (ord("H"), ord("O"))]
        assert n_ev[c] == len(votes)
        if not votes:
            continue
        called, q = "".join(v[0] for v in votes), [v[1] for v in votes]
        a = orc.consensus_error_rate(P, pdf, called, q, [with_prior[chr(ref[c])][b] for b in cs.BASES])
        b = orc.consensus_error_rate(P, pdf, called, q, [without_prior[chr(ref[c])][b] for b in cs.BASES])
        assert same(a, [float(x) for x in p_w[c]]) and same(b, [float(x) for x in p_wo[c]]), c
        special += any(math.isnan(x) or math.isinf(x) for x in a)
    assert special > 50, "the columns must reach the NaN / infinity paths"
