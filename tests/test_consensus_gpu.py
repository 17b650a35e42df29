"""Row N3, second half, on the device: consensus.calculate_consensus (smx_consensus_run on top of msa.generate_msa) ==
the UNMODIFIED reference's MyMSA.calculate_consensus on the golden cases (tests/golden/consensus_v1.json.gz): both
consensus triples and the recorded p_lists bit for bit, from host matrices and from the matrices smx_msa_run left on the
device; what the reference crashes on is an error; and a full-depth stack (3 000 reads x 8 000 columns) against the CPU
oracle on sampled columns, with the cells/s figure printed."""
import gzip
import json
import time
from pathlib import Path

import numpy as np
import pytest

from oracle import oracle as orc

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).parent / "golden" / "consensus_v1.json.gz"


def gold():
    with gzip.open(GOLD, "rt") as f:
        return json.load(f)


def test_consensus_matches_reference(engine):
    from multiplexnanopore_b200 import consensus as cs
    from multiplexnanopore_b200.msa import generate_msa
    g = gold()
    pdfs = [cs.SequenceBasecallQscorePDF(columns=t["columns"], counts=t["counts"]) for t in g["tables"]]
    for case in g["cases"]:
        params = dict(case["params"])
        for resident in (False, True):
            m = generate_msa(case["ref"], case["queries"], case["q_scores"], case["cigars"], params, engine=engine)
            p_w, p_wo, n_ev = cs.consensus_error_rates(m, pdfs[case["table"]], engine=engine, device_resident=resident)
            for c, rec in enumerate(case["p_lists"]):
                if rec is None:
                    assert n_ev[c] == 0, (case["name"], c)
                elif rec:
                    assert [float(x).hex() for x in p_w[c]] == rec[0] and [float(x).hex() for x in p_wo[c]] == rec[1], (case["name"], c, resident)
            w, wo = cs.calculate_consensus(m, pdfs[case["table"]], engine=engine, device_resident=False)
            assert [w[0], [int(x) for x in w[1]], w[2]] == case["with_prior"], case["name"]
            assert [wo[0], [int(x) for x in wo[1]], wo[2]] == case["without_prior"], case["name"]
            assert m.with_prior_consensus_seq == case["with_prior"][0] and m.without_prior_consensus_my_cigar == case["without_prior"][2]


def test_consensus_rejects_what_crashes_the_reference(engine):
    from multiplexnanopore_b200 import consensus as cs
    from multiplexnanopore_b200.msa import generate_msa
    t = gold()["tables"][0]
    pdf = cs.SequenceBasecallQscorePDF(columns=t["columns"], counts=t["counts"])
    params = dict(error_rate=1e-7, ins_rate=1e-7)
    m = generate_msa("ACGT", ["ACNT"], [[30] * 4], ["===="], params, engine=engine)     # an N call: missing key in the reference's C++ maps
    with pytest.raises(ValueError):
        cs.calculate_consensus(m, pdf, engine=engine)
    m = generate_msa("ACGT", ["ACGT"], [[30] * 4], ["===="], params, engine=engine)
    with pytest.raises(Exception):   # dimensions of the resident matrices do not match
        generate_msa("ACGTA", ["ACGTA"], [[30] * 5], ["====="], params, engine=engine)
        cs.consensus_error_rates(m, pdf, engine=engine, device_resident=True)


def test_full_depth_stack(engine):
    from multiplexnanopore_b200 import consensus as cs
    from multiplexnanopore_b200.msa import MyMSA
    rng = np.random.default_rng(41)
    t = gold()["tables"][0]
    pdf = cs.SequenceBasecallQscorePDF(columns=t["columns"], counts=t["counts"])
    R, C = 3000, 8000
    ref = rng.choice(np.frombuffer(b"ACGT", np.uint8), size=C)
    seq = np.repeat(ref[None, :], R, axis=0)
    u = rng.random((R, C))
    sub = u < 0.04
    seq[sub] = rng.choice(np.frombuffer(b"ACGT", np.uint8), size=int(sub.sum()))
    dele = (u >= 0.04) & (u < 0.07)
    seq[dele] = ord("-")
    qs = rng.integers(1, 42, size=(R, C)).astype(np.int8)
    qs[dele] = -1
    cig = np.full((R, C), ord("="), np.uint8)
    cig[dele] = ord("D")
    edge = rng.random((R, C)) < 0.02
    cig[edge] = ord("H"); seq[edge] = ord(" "); qs[edge] = -1
    params = dict(error_rate=1e-7, ins_rate=1e-7)
    m = MyMSA(ref, seq, qs, cig, params)
    cs.consensus_error_rates(m, pdf, engine=engine)   # warm-up
    t0 = time.perf_counter()
    p_w, p_wo, n_ev = cs.consensus_error_rates(m, pdf, engine=engine)
    dt = time.perf_counter() - t0
    print(f"\nconsensus: {R} reads x {C} columns = {R * C / 1e6:.0f} M cells (x 25 chains x 2 priors) in {dt * 1e3:.0f} ms = {R * C / dt / 1e9:.2f} G cells/s incl. upload and host finish")
    P, pd = pdf.tables()
    with_prior, without_prior = cs.consensus_params(params)
    for c in rng.choice(C, size=40, replace=False):
        votes = [(chr(seq[r, c]), int(qs[r, c])) for r in range(R) if cig[r, c] not in (ord("H"), ord("O"))]
        assert n_ev[c] == len(votes)
        called = "".join(v[0] for v in votes); q = [v[1] for v in votes]
        r = chr(ref[c])
        assert orc.consensus_error_rate(P, pd, called, q, [with_prior[r][b] for b in cs.BASES]) == [float(x) for x in p_w[c]]
        assert orc.consensus_error_rate(P, pd, called, q, [without_prior[r][b] for b in cs.BASES]) == [float(x) for x in p_wo[c]]
    w, wo = cs.calculate_consensus(m, pdf, engine=engine)
    assert w[0].count("-") < C // 50 and sum(a == chr(b) for a, b in zip(w[0], ref)) > 0.99 * C
