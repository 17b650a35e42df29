"""Edge inputs through the reference-facing API (execute_alignment_core, calc_distance) against the oracle port:
empty FASTQ, one read, reads too short for any conserved region, all-N reads, lower-case input, reads longer than
1.75 references (omit_too_long), a read batch boundary (1001 reads -> two batches in flight)."""
import numpy as np
import pytest

from multiplexnanopore_b200 import savemoney as sm
from multiplexnanopore_b200 import synth
from oracle import oracle as orc
from oracle import savemoney_port as port

pytestmark = pytest.mark.gpu
PARAMS = dict(gap_open_penalty=3, gap_extend_penalty=1, match_score=1, mismatch_score=-2, topology_of_dna=0)


def _want(plasmids, read, params, omit=True):
    return [[orc.rle(r.my_cigar), int(r.score), None if r.new_query_seq_offset is None else int(r.new_query_seq_offset)]
            for r in port.align_read([p.upper() for p in plasmids], read.upper(), params, omit)]


def test_edge_reads(engine):
    rng = np.random.default_rng(81)
    plasmids = synth.plasmid_set(rng, 3, 900, 1600, backbone=400)
    good, _ = synth.reads_for(rng, plasmids, 2, topology=0)
    reads = {
        "one_base": "A", "too_short": plasmids[0][100:118], "just_21": plasmids[1][200:221], "all_n": "N" * 300,
        "lower": good[0].lower(), "long": (good[1] * 3)[: int(len(plasmids[0]) * 2.2)], "good": good[1],
        "poly_a": "A" * 500,
    }
    fastq = {k: [v, [20] * len(v)] for k, v in reads.items()}
    assert sm.execute_alignment_core(plasmids, {}, PARAMS, engine=engine) == {}
    rd = sm.execute_alignment_core([p.lower() for p in plasmids], fastq, PARAMS, engine=engine)
    assert list(rd) == list(reads)
    for qid, read in reads.items():
        got = [[r.cigar, r.score, r.new_query_seq_offset] for r in rd[qid]]
        assert got == _want(plasmids, read, PARAMS), qid
    assert all(r.my_cigar == "" and r.score == 0 and r.new_query_seq_offset is None for r in rd["one_base"])
    one = sm.execute_alignment_core(plasmids, {"good": fastq["good"]}, PARAMS, engine=engine)
    assert [[r.cigar, r.score, r.new_query_seq_offset] for r in one["good"]] == _want(plasmids, reads["good"], PARAMS)


def test_batch_boundary_and_order(engine):
    """1001 short reads: two batches (1000 + 1) in flight; results must come back in read order."""
    rng = np.random.default_rng(82)
    plasmids = synth.plasmid_set(rng, 2, 300, 420, backbone=120)
    reads, _ = synth.reads_for(rng, plasmids, 1001, topology=0, frac_full=0.3, frac_partial=0.6)
    res = sm.align_all(plasmids, reads, PARAMS, True, engine=engine)
    n2 = 2 * len(plasmids)
    assert len(res.score) == 1001 * n2 and int(res.stats["n_pair_strands"]) == 1001 * n2
    for qi in list(range(0, 1001, 97)) + [999, 1000]:
        want = _want(plasmids, reads[qi], PARAMS)
        got = [[res.cigar_string(qi * n2 + j), int(res.score[qi * n2 + j]), res.offset(qi * n2 + j)] for j in range(n2)]
        assert got == want, qi


def test_calc_distance_edges(engine):
    rng = np.random.default_rng(83)
    a = synth.random_bases(rng, 700).tobytes().decode()
    b = synth.random_bases(rng, 650).tobytes().decode()
    params = dict(PARAMS)
    assert sm.calc_distance(a, a, params, engine=engine) == 0
    assert sm.calc_distance(a, synth.revcomp_bytes(np.frombuffer(a.encode(), np.uint8)).tobytes().decode(), params, engine=engine) == 0
    assert sm.calc_distance(a, b, params, engine=engine) == port.distance(a, b, params)
    assert sm.set_distance_matrix([a], params, engine=engine).tolist() == [[0]]


def test_ultra_long_reads_do_not_fail_the_call(engine):
    """a read of more than 65 535 bases made smx_pipeline_run fail for the whole FASTQ.  Such a read is
    (1) dropped by omit_too_long like in the reference (empty MyResult(), post_analysis_core.py:76-83) when it is longer
    than 1.75 references, and (2) aligned normally when it is not: its pair-strands take the host re-decision of the
    threshold (the select kernel's histogram median holds 16-bit counts).  The other reads of the batch are unaffected."""
    rng = np.random.default_rng(84)
    small = synth.plasmid_set(rng, 2, 900, 1300, backbone=300)
    good, _ = synth.reads_for(rng, small, 3, topology=0)
    monster = synth.random_bases(rng, 70000).tobytes().decode()
    fastq = {"good0": [good[0], None], "monster": [monster, None], "good1": [good[1], None]}
    rd = sm.execute_alignment_core(small, fastq, PARAMS, engine=engine)
    assert all(r.my_cigar == "" and r.score == 0 for r in rd["monster"])
    for k, read in (("good0", good[0]), ("good1", good[1])):
        assert [[r.cigar, r.score, r.new_query_seq_offset] for r in rd[k]] == _want(small, read, PARAMS), k
    # not omitted: a 66 kb read against a 38 kb plasmid it was copied from (rotated, 1.73 x its length, 0.4 % errors so
    # that the CPU port's cubic chain search stays at seconds)
    big = synth.random_bases(rng, 38000).tobytes().decode()
    src = (big[15000:] + big[:15000]) * 2
    long_read = synth.mutate_bytes(rng, np.frombuffer(src[:65900].encode(), np.uint8), 0.002, 0.001, 0.001).tobytes().decode()
    assert 65535 < len(long_read) < 1.75 * len(big)
    res = sm.align_all([big], [long_read, good[2]], PARAMS, True, engine=engine)
    assert int(res.stats["n_host_redecisions"]) >= 2 and int(res.stats["n_failed"]) == 0
    want = _want([big], long_read, PARAMS)
    got = [[res.cigar_string(j), int(res.score[j]), res.offset(j)] for j in range(2)]
    assert got == want
    assert max(w[1] for w in want) > 30000   # it did align


def test_mid_size_call_takes_the_adaptive_batches(engine):
    """1 000 < reads < batch x overlap: the call is cut into smaller batches so that more handles are busy (one GPU's share
    of a strong-scaled run); results and order are those of one big batch."""
    rng = np.random.default_rng(85)
    plasmids = synth.plasmid_set(rng, 2, 300, 420, backbone=120)
    reads, _ = synth.reads_for(rng, plasmids, 1300, topology=0, frac_full=0.3, frac_partial=0.6)
    res = sm.align_all(plasmids, reads, PARAMS, True, engine=engine)          # 1300 reads -> 3 batches of 500
    assert len(getattr(res, "parts", [])) >= 3
    one = sm.align_all(plasmids, reads, PARAMS, True, engine=engine, batch_queries=2000)   # one pipeline call
    assert np.array_equal(res.score, one.score) and np.array_equal(res.status, one.status) and np.array_equal(res.cigar_off, one.cigar_off)
    assert all(np.array_equal(res.runs(i), one.runs(i)) for i in range(0, len(res.score), 37))
