"""Randomized parity campaign: the CUDA engine against the CPU oracle port on inputs that stress the corners --
IUPAC letters in plasmid maps, N in reads, tandem repeats (overlapping regions in the query dimension), very short
and very long reads (omit_too_long), five scorings, circular and linear topology, omit_too_long on and off.
(The same generator ran for 4 000 reads during development: scratch/fuzz.py.)"""
import multiprocessing as mp
import os

import numpy as np
import pytest

from multiplexnanopore_b200 import savemoney as sm
from multiplexnanopore_b200 import synth
from oracle import oracle as orc
from oracle import savemoney_port as port

pytestmark = pytest.mark.gpu


def _port_read(args):
    plasmids, read, params, omit = args
    try:
        return [(orc.rle(r.my_cigar), int(r.score), None if r.new_query_seq_offset is None else int(r.new_query_seq_offset))
                for r in port.align_read(plasmids, read, params, omit)]
    except AssertionError:
        return "ASSERT"


def _sprinkle(rng, s, letters, rate):
    a = np.frombuffer(s.encode(), dtype=np.uint8).copy()
    m = rng.random(a.size) < rate
    a[m] = np.frombuffer(letters.encode(), dtype=np.uint8)[rng.integers(0, len(letters), size=int(m.sum()))]
    return a.tobytes().decode()


def make_case(case, seed0=9000):
    rng = np.random.default_rng(seed0 + case)
    topo = int(rng.random() < 0.35)
    npl = int(rng.integers(2, 5))
    plasmids = synth.plasmid_set(rng, npl, 700, 2600, backbone=int(rng.integers(200, 700)))
    kind = case % 6
    if kind == 1:
        plasmids = [_sprinkle(rng, p, "NRYKMSW", 0.003) for p in plasmids]
    reads, _ = synth.reads_for(rng, plasmids, 10, topology=topo, frac_full=0.5, frac_partial=0.3,
                               homopolymer_rate=float(rng.choice([0.0, 1.0, 4.0])), sub=float(rng.choice([0.01, 0.03, 0.08])),
                               ins=float(rng.choice([0.005, 0.025, 0.06])), dele=float(rng.choice([0.005, 0.025, 0.06])))
    if kind == 2:
        reads = [_sprinkle(rng, r, "N", 0.004) for r in reads]
    if kind == 3:
        p0 = plasmids[0]
        a = int(rng.integers(100, len(p0) - 400))
        plasmids[0] = p0[:a + 300] + p0[a:a + 300] + p0[a + 300:]
        reads, _ = synth.reads_for(rng, plasmids, 10, topology=topo, frac_full=0.7, frac_partial=0.2)
    if kind == 4:
        reads = reads[:5] + [reads[5][:int(rng.integers(22, 120))], reads[6][:int(rng.integers(120, 400))],
                             (reads[7] * 3)[:int(len(plasmids[0]) * 2.4)], reads[8], reads[9][:25]]
    sc = [(3, 1, 1, -2), (5, 2, 2, -3), (2, 2, 1, -1), (4, 1, 3, -2), (3, 1, 1, -2)][case % 5]
    params = dict(gap_open_penalty=sc[0], gap_extend_penalty=sc[1], match_score=sc[2], mismatch_score=sc[3], topology_of_dna=topo)
    return plasmids, reads, params, bool(case % 7 != 3)


def test_randomized_cases_match_port(engine):
    orc.build()
    cases = [make_case(c) for c in range(36)]
    with mp.get_context("fork").Pool(processes=min(16, os.cpu_count() or 1)) as pool:
        want = pool.map(_port_read, [(pl, r, pa, om) for pl, rd, pa, om in cases for r in rd], chunksize=2)
    k = 0
    for ci, (plasmids, reads, params, omit) in enumerate(cases):
        res = sm.align_all(plasmids, reads, params, omit, engine=engine)
        n2 = 2 * len(plasmids)
        for qi in range(len(reads)):
            w = want[k]; k += 1
            if w == "ASSERT":
                assert (res.status[qi * n2:(qi + 1) * n2] == 2).any()
                continue
            got = [(res.cigar_string(qi * n2 + j), int(res.score[qi * n2 + j]), res.offset(qi * n2 + j)) for j in range(n2)]
            assert got == w, f"case {ci} read {qi} params {params}"
