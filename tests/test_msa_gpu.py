"""Row N3 on the device: msa.generate_msa (smx_msa_run) == the UNMODIFIED reference's MyMSA.generate_msa on the golden
cases (tests/golden/msa_v1.json.gz), fed with column strings and with run arrays; its errors; and a full-size stack
(3 000 reads on an 8 kb plasmid) checked through column invariants, with the columns/s figure printed."""
import gzip
import json
import time
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).parent / "golden" / "msa_v1.json.gz"


def test_generate_msa_matches_reference(engine):
    from multiplexnanopore_b200.msa import cigar_runs, generate_msa
    with gzip.open(GOLD, "rt") as f:
        cases = json.load(f)["cases"]
    for c in cases:
        for as_runs in (False, True):
            cigs = [cigar_runs(x) for x in c["cigars"]] if as_runs else c["cigars"]
            m = generate_msa(c["ref"], c["queries"], c["q_scores"], cigs, engine=engine)
            assert m.ref_seq_aligned == c["ref_seq_aligned"], c["name"]
            assert m.query_seq_list_aligned == c["query_seq_list_aligned"], c["name"]
            assert m.q_scores_list_aligned == c["q_scores_list_aligned"], c["name"]
            assert m.my_cigar_list_aligned == c["my_cigar_list_aligned"], c["name"]


def test_generate_msa_rejects_what_the_reference_rejects(engine):
    from multiplexnanopore_b200.msa import generate_msa
    with pytest.raises(AssertionError):   # CIGAR shorter than the reference: the final assertions of msa.py:913-916
        generate_msa("ACGTACGT", ["ACGT"], [[30] * 4], ["===="], engine=engine)
    with pytest.raises(AssertionError):   # query longer than the CIGAR consumes
        generate_msa("ACGT", ["ACGTA"], [[30] * 5], ["===="], engine=engine)
    with pytest.raises(Exception):        # a letter generate_msa does not know in a reference column
        generate_msa("ACGT", ["ACGT"], [[30] * 4], ["==N="], engine=engine)
    with pytest.raises(AssertionError):
        generate_msa("ACGT", ["ACGT"], [[30] * 4], [], engine=engine)


def test_full_size_stack(engine):
    from multiplexnanopore_b200 import savemoney as sm, synth
    from multiplexnanopore_b200.msa import generate_msa
    rng = np.random.default_rng(33)
    plasmid = synth.plasmid_set(rng, 1, 8000, 8000, backbone=2500)[0]
    reads, _ = synth.reads_for(rng, [plasmid], 3000, topology=0)
    p = dict(gap_open_penalty=3, gap_extend_penalty=1, match_score=1, mismatch_score=-2, topology_of_dna=0)
    res = sm.align_all([plasmid], reads, p, True, engine=engine)
    comp = str.maketrans("ACGT", "TGCA")
    qs, cigs, scores = [], [], []
    for k, read in enumerate(reads):   # what MyMSAligner.execute feeds (msa.py:617-629): the better strand, rotated by its offset
        j = 2 * k if res.score[2 * k] >= res.score[2 * k + 1] else 2 * k + 1
        if res.status[j] != 0 or res.score[j] < 0.3 * len(plasmid):
            continue
        seq = read if j % 2 == 0 else read.translate(comp)[::-1]
        off = res.offset(j)
        qs.append(seq[off:] + seq[:off]); cigs.append(res.runs(j)); scores.append(rng.integers(1, 42, size=len(seq)).astype(np.int8))
    assert len(qs) > 2000
    generate_msa(plasmid, qs[:50], scores[:50], cigs[:50], engine=engine)   # warm-up
    t0 = time.perf_counter()
    m = generate_msa(plasmid, qs, scores, cigs, engine=engine)
    dt = time.perf_counter() - t0
    n, c = m.my_cigar_matrix.shape
    print(f"\ngenerate_msa: {n} reads x {len(plasmid)} bases -> {c} columns in {dt * 1e3:.0f} ms = {n * c / dt / 1e9:.2f} G cells/s (host packing and the copy back included)")
    cig, seq = m.my_cigar_matrix, m.query_seq_matrix
    ref_al = m.ref_seq_aligned_matrix
    assert (ref_al != ord("-")).sum() == len(plasmid) and bytes(ref_al[ref_al != ord("-")]).decode() == plasmid
    extra = ref_al == ord("-")
    # extra columns hold only I / S / fillers, reference columns never do; every extra column has at least one consumer
    assert np.isin(cig[:, extra], np.frombuffer(b"ISNO", np.uint8)).all() and not np.isin(cig[:, ~extra], np.frombuffer(b"ISNO", np.uint8)).any()
    assert np.isin(cig[:, extra], np.frombuffer(b"IS", np.uint8)).any(axis=0).all()
    for k in range(0, n, 97):   # a row without its fillers is the read and its CIGAR again
        row_c = bytes(cig[k][~np.isin(cig[k], np.frombuffer(b"NO", np.uint8))]).decode()
        from multiplexnanopore_b200.realign import expand_all
        want = expand_all(np.array([0, cigs[k].size], np.int64), cigs[k])[0]
        assert row_c == want
        has_base = np.isin(cig[k], np.frombuffer(b"=XIS", np.uint8))
        assert bytes(seq[k][has_base]).decode() == qs[k]
        assert (m.q_scores_matrix[k][has_base] == scores[k]).all() and (m.q_scores_matrix[k][~has_base] == -1).all()
