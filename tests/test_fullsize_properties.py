"""BASELINE.json full-size configurations through size-independent properties (the oracle cannot
finish these sizes): every final CIGAR consumes exactly N_ref reference and N_query query bases, the
score equals the affine score of the CIGAR (set_cigar_score, ref_query_alignment.py:452-464), '=' / 'X'
columns are right (AlignmentBase.assert_alignment, my_classes.py:407-423), results are deterministic,
reads are assigned to the plasmid/strand they were drawn from, pre_survey matrices are symmetric with
a zero diagonal and recover the family structure."""
import numpy as np
import pytest

from multiplexnanopore_b200 import savemoney as sm
from multiplexnanopore_b200 import synth

pytestmark = pytest.mark.gpu

PARAMS = dict(gap_open_penalty=3, gap_extend_penalty=1, match_score=1, mismatch_score=-2, topology_of_dna=0)
_COMP = str.maketrans("ACGTMRWSYKVHDBN", "TGCAKYWSRMBDHVN")


def run_sums(res):
    """per pair-strand sums of run lengths by op -> dict op -> int64[n]"""
    n = res.score.size
    ops = res.cigar_rle & 15
    lens = (res.cigar_rle >> 4).astype(np.int64)
    owner = np.repeat(np.arange(n), np.diff(res.cigar_off))
    return {o: np.bincount(owner[ops == o], weights=lens[ops == o], minlength=n).astype(np.int64) for o in (1, 2, 4, 5, 7, 8)}, owner, ops, lens


def check_structure(res, ref_lens, qry_lens, n_refs, params):
    go, ge, ma, mi = (params[k] for k in ("gap_open_penalty", "gap_extend_penalty", "match_score", "mismatch_score"))
    n = res.score.size
    s, owner, ops, lens = run_sums(res)
    nonempty = res.status == 0
    assert (res.status != 2).all()
    pair = np.arange(n) // 2
    nr = np.asarray(ref_lens)[pair % n_refs]
    nq = np.asarray(qry_lens)[pair // n_refs]
    ref_used = s[7] + s[8] + s[2] + s[5]
    qry_used = s[7] + s[8] + s[1] + s[4]
    assert np.array_equal(ref_used[nonempty], nr[nonempty])
    assert np.array_equal(qry_used[nonempty], nq[nonempty])
    assert (np.diff(res.cigar_off)[~nonempty] == 0).all() and (res.score[~nonempty] == 0).all()
    # score = sum over maximal runs (adjacent runs of one pair-strand never share a letter)
    same = (owner[1:] == owner[:-1]) & (ops[1:] == ops[:-1])
    assert not same.any()
    contrib = np.where(ops == 7, ma * lens, 0) + np.where(ops == 8, mi * lens, 0) - np.where((ops == 1) | (ops == 2), go + ge * (lens - 1), 0)
    score = np.bincount(owner, weights=contrib, minlength=n).astype(np.int64)
    assert np.array_equal(score, res.score.astype(np.int64))
    off = res.new_query_seq_offset
    assert (off[~nonempty] == -2147483648).all()


def check_columns(res, refs, reads, n_refs, idxs):
    for i in idxs:
        cig = res.my_cigar(int(i))
        p, strand = int(i) // 2, int(i) % 2
        ref, q = refs[p % n_refs], reads[p // n_refs]
        if strand:
            q = q.translate(_COMP)[::-1]
        o = res.offset(int(i))
        q = q[o:] + q[:o]
        r = qq = 0
        for c in cig:
            if c in "=X":
                assert (ref[r] == q[qq]) == (c == "="), f"pair-strand {i}"
                r += 1; qq += 1
            elif c in "IS":
                qq += 1
            else:
                r += 1
        assert (r, qq) == (len(ref), len(q))


def test_c2_full_size(engine):
    plasmids, reads, truth = synth.config_c2(n_reads=20000)
    res = sm.align_all(plasmids, reads, PARAMS, True, engine=engine)
    n_refs = len(plasmids)
    check_structure(res, [len(p) for p in plasmids], [len(r) for r in reads], n_refs, PARAMS)
    rng = np.random.default_rng(0)
    aligned = np.where(res.status == 0)[0]
    check_columns(res, plasmids, reads, n_refs, rng.choice(aligned, size=150, replace=False))
    # read assignment (msa.QueryAssignment, msa.py:31-64): normalised score, unique argmax, threshold 0.3
    sc = res.score.reshape(len(reads), 2 * n_refs).astype(float)
    norm = sc / np.repeat([len(p) for p in plasmids], 2)[None, :]
    best = norm.argmax(1)
    expected = np.array([2 * pi + int(rc) for pi, rc in truth])
    full = np.array([abs(len(r) - len(plasmids[pi])) < 0.2 * len(plasmids[pi]) for r, (pi, _) in zip(reads, truth)])
    assert (best[full] == expected[full]).mean() > 0.995
    assert (norm.max(1)[full] > 0.3).mean() > 0.99
    # determinism: same batch again
    res2 = sm.align_all(plasmids, reads[:3000], PARAMS, True, engine=engine)
    k = 3000 * 2 * n_refs
    assert np.array_equal(res2.score, res.score[:k]) and np.array_equal(res2.new_query_seq_offset, res.new_query_seq_offset[:k])
    assert np.array_equal(res2.cigar_rle, res.cigar_rle[: res.cigar_off[k]])


def test_c3_pre_survey_matrix(engine):
    maps = synth.config_c3(seed=3, n_maps=48, families=8, len_lo=2000, len_hi=15000)
    dm = sm.set_distance_matrix(maps, PARAMS, engine=engine)
    assert (dm == dm.T).all() and (np.diag(dm) == 0).all() and (dm >= 0).all()
    fam = np.arange(48) // 6
    same = fam[:, None] == fam[None, :]
    off = ~np.eye(48, dtype=bool)
    assert dm[same & off].max() < 400          # <= 50 edits each -> small distances inside a family
    assert dm[~same].min() > 1500              # unrelated maps: len_i + len_j or a poor alignment


def test_c4_linear_subset(engine):
    rng = np.random.default_rng(4)
    plasmids = synth.plasmid_set(rng, 12, 3000, 20000, backbone=2000)
    reads, truth = synth.reads_for(rng, plasmids, 1500, topology=1)
    params = dict(PARAMS, topology_of_dna=1)
    res = sm.align_all(plasmids, reads, params, True, engine=engine)
    check_structure(res, [len(p) for p in plasmids], [len(r) for r in reads], len(plasmids), params)
    assert (res.new_query_seq_offset[res.status == 0] == 0).all()
    aligned = np.where(res.status == 0)[0]
    check_columns(res, plasmids, reads, len(plasmids), rng.choice(aligned, size=60, replace=False))
    sc = res.score.reshape(len(reads), 24).astype(float) / np.repeat([len(p) for p in plasmids], 2)[None, :]
    expected = np.array([2 * pi + int(rc) for pi, rc in truth])
    assert (sc.argmax(1) == expected).mean() > 0.99


@pytest.mark.parametrize("topology", [0, 1])
def test_device_chain_equals_host_chain(engine, topology, monkeypatch):
    """row A6 has two independent implementations (smx_chain.cuh on the device, smx_chain.hpp on the host,
    selected by SMX_CHAIN=host); on a few thousand pair-strands every output must be identical."""
    rng = np.random.default_rng(77 + topology)
    plasmids = synth.plasmid_set(rng, 6, 4000, 9000, backbone=2500)
    reads, _ = synth.reads_for(rng, plasmids, 600, topology=topology, homopolymer_rate=0.5)
    params = dict(PARAMS, topology_of_dna=topology)
    monkeypatch.delenv("SMX_CHAIN", raising=False)
    dev = sm.align_all(plasmids, reads, params, True, engine=engine)
    monkeypatch.setenv("SMX_CHAIN", "host")
    host = sm.align_all(plasmids, reads, params, True, engine=engine)
    assert np.array_equal(dev.status, host.status)
    assert np.array_equal(dev.score, host.score)
    assert np.array_equal(dev.new_query_seq_offset, host.new_query_seq_offset)
    assert np.array_equal(dev.cigar_off, host.cigar_off) and np.array_equal(dev.cigar_rle, host.cigar_rle)


@pytest.mark.parametrize("topology", [0, 1])
def test_device_select_equals_host_redecision(engine, topology, monkeypatch):
    """Row A4 has two implementations as well: the select kernel and the host re-decision that takes over when a
    threshold lies within 1e-6 of an integer or a pair-strand overflows the device slots (forced for every
    pair-strand by SMX_SELECT=host: numpy's exact summation order); every output must be identical."""
    rng = np.random.default_rng(97 + topology)
    plasmids = synth.plasmid_set(rng, 4, 1500, 4000, backbone=800)
    reads, _ = synth.reads_for(rng, plasmids, 60, topology=topology)
    params = dict(PARAMS, topology_of_dna=topology)
    monkeypatch.delenv("SMX_SELECT", raising=False)
    dev = sm.align_all(plasmids, reads, params, True, engine=engine)
    monkeypatch.setenv("SMX_SELECT", "host")
    host = sm.align_all(plasmids, reads, params, True, engine=engine)
    assert host.stats["n_host_redecisions"] > 0 and dev.stats["n_host_redecisions"] == 0
    assert np.array_equal(dev.score, host.score) and np.array_equal(dev.status, host.status)
    assert np.array_equal(dev.new_query_seq_offset, host.new_query_seq_offset)
    assert np.array_equal(dev.cigar_off, host.cigar_off) and np.array_equal(dev.cigar_rle, host.cigar_rle)


def test_run_length_merge_equals_column_string_merge(engine, monkeypatch):
    """Row A10 has two implementations: the run-length overlap merge (default) and the column-string one
    (SMX_MERGE=chars, the first implementation; tests/test_stitch_cpu.py compares them on random rows without a GPU);
    on real pipeline output every pair-strand must be identical."""
    rng = np.random.default_rng(131)
    plasmids = synth.plasmid_set(rng, 6, 4000, 9000, backbone=2500)
    reads, _ = synth.reads_for(rng, plasmids, 600, topology=0, homopolymer_rate=0.5)
    monkeypatch.delenv("SMX_MERGE", raising=False)
    runs = sm.align_all(plasmids, reads, PARAMS, True, engine=engine)
    monkeypatch.setenv("SMX_MERGE", "chars")
    chars = sm.align_all(plasmids, reads, PARAMS, True, engine=engine)
    assert int(runs.stats.get("n_failed", 0)) == int(chars.stats.get("n_failed", 0))
    assert np.array_equal(runs.status, chars.status) and np.array_equal(runs.score, chars.score)
    assert np.array_equal(runs.new_query_seq_offset, chars.new_query_seq_offset)
    assert np.array_equal(runs.cigar_off, chars.cigar_off) and np.array_equal(runs.cigar_rle, chars.cigar_rle)
