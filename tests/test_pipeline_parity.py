"""Rows A1-A12 end to end: CUDA engine (scan -> select -> chain -> DP -> stitch) through the C-ABI
== reference outputs recorded in tests/golden/ and == the CPU oracle port on fresh seeded inputs."""
import numpy as np
import pytest

from multiplexnanopore_b200 import savemoney as sm
from multiplexnanopore_b200 import synth
from oracle import oracle as orc
from oracle import savemoney_port as port
from tests import golden_util as gu

pytestmark = pytest.mark.gpu


def triples(res, i):
    return [res.cigar_string(i), int(res.score[i]), res.offset(i)]


@pytest.mark.parametrize("name", gu.CASE_NAMES)
def test_engine_matches_reference_golden(engine, name):
    c = gu.case(name)
    res = sm.align_all(c["refs"], c["reads"], c["params"], c["omit_too_long"], engine=engine)
    assert (res.status != 2).all()
    n2 = 2 * len(c["refs"])
    for qi, want in enumerate(c["results"]):
        got = [triples(res, qi * n2 + j) for j in range(n2)]
        assert got == want, f"{name}: read {qi}"
    # DP cell accounting identical to the reference's parasail call list
    assert int(res.stats["n_dp_problems"]) == len(c["dp_calls"])
    assert int(res.stats["dp_cells"]) == sum(a * b for _, a, b in c["dp_calls"])


def test_engine_matches_golden_distances(engine):
    g = gu.load()
    names = g["distances"]["demo_names"]
    seqs = [g["demo_maps"][n] for n in names]
    params = g["cases"][0]["params"]
    dm = sm.set_distance_matrix(seqs, params, engine=engine)
    for key, want in g["distances"]["demo"].items():
        i, j = map(int, key.split(","))
        assert dm[i, j] == want == dm[j, i]
    fam = g["distances"]["family_maps"]
    dm = sm.set_distance_matrix(fam, params, engine=engine)
    for key, want in g["distances"]["family"].items():
        i, j = map(int, key.split(","))
        assert dm[i, j] == want


@pytest.mark.parametrize("topology,seed", [(0, 31), (0, 32), (1, 33)])
def test_engine_matches_port_on_fresh_inputs(engine, topology, seed):
    rng = np.random.default_rng(seed)
    plasmids = synth.plasmid_set(rng, 4, 1200, 2600, backbone=600)
    reads, _ = synth.reads_for(rng, plasmids, 10, topology=topology, frac_full=0.6, frac_partial=0.25, homopolymer_rate=1.0)
    params = dict(gap_open_penalty=3, gap_extend_penalty=1, match_score=1, mismatch_score=-2, topology_of_dna=topology)
    res = sm.align_all(plasmids, reads, params, True, engine=engine)
    n2 = 2 * len(plasmids)
    for qi, read in enumerate(reads):
        want = port.align_read(plasmids, read, params, True)
        for j, w in enumerate(want):
            i = qi * n2 + j
            assert triples(res, i) == [orc.rle(w.my_cigar), int(w.score), None if w.new_query_seq_offset is None else int(w.new_query_seq_offset)], f"read {qi} pair-strand {j}"


def test_execute_alignment_core_interface(engine):
    c = gu.case("c2_small_circular")
    fastq = {f"@read{i}": [r, [30] * len(r)] for i, r in enumerate(c["reads"][:4])}
    rd = sm.execute_alignment_core(c["refs"], fastq, c["params"], engine=engine)
    assert list(rd.keys()) == list(fastq.keys())
    for qi, (qid, row) in enumerate(rd.items()):
        assert len(row) == 2 * len(c["refs"])
        for r, want in zip(row, c["results"][qi]):
            assert [r.to_dict()["cigar"], r.score, r.new_query_seq_offset] == want
            assert r.cigar == want[0] and len(r) == sum(int(x) for x in __import__("re").findall(r"(\d+)", want[0]))
