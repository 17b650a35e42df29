"""Row N4: the legacy dense mode (sw_trace(read, ref + ref) per read x plasmid x strand,
Modules/1_alignment_consensus_core.py:219-296) through savemoney.legacy_dense_align == the oracle's sw_trace."""
import numpy as np
import pytest

from oracle import oracle as orc
from tests import util

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("topology", [0, 1])
def test_legacy_dense_matches_oracle(engine, topology):
    from multiplexnanopore_b200 import savemoney as sm
    rng = np.random.default_rng(71 + topology)
    refs = [util.rand_seq(rng, int(rng.integers(300, 900))) for _ in range(3)]
    fastq = {}
    for k in range(7):
        src = refs[int(rng.integers(3))]
        a = int(rng.integers(len(src)))
        read = util.mutate(rng, src[a:] + src[:a])[: int(rng.integers(60, 1000))]
        if rng.random() < 0.5:
            read = read.translate(str.maketrans("ACGT", "TGCA"))[::-1]
        fastq[f"r{k}"] = [read, [30] * len(read)]
    params = dict(gap_open_penalty=3, gap_extend_penalty=1, match_score=1, mismatch_score=-2, topology_of_dna=topology)
    got = sm.legacy_dense_align(refs, fastq, params, engine=engine, batch_cells=4_000_000)
    comp = str.maketrans("ACGT", "TGCA")
    for qid, (read, _) in fastq.items():
        assert len(got[qid]) == 2 * len(refs)
        for t, ref in enumerate(refs):
            target = ref + ref if topology == 0 else ref
            for strand, q in enumerate((read, read.translate(comp)[::-1])):
                o = orc.align(3, q, target, 3, 1, 1, -2)
                g = got[qid][2 * t + strand]
                assert (g.score, g.cigar, g.beg_query, g.beg_ref, g.end_query, g.end_ref) == \
                    (o.score, o.cigar_rle, o.beg_query, o.beg_ref, o.end_query, o.end_ref), (qid, t, strand)


def test_legacy_dense_c2_sized_reads(engine):
    """reads of 5-8 kb against doubled 5-7 kb plasmids (the packed local kernel incl. its teams of 2 / 4 warps), a read
    that matches both copies of the doubled reference equally well (end-cell tie: smallest reference position wins), and
    the leading-'D' quirk of sw_trace the reference documents (ref_query_alignment.py:37-54): whatever the CIGAR starts
    with, engine == oracle."""
    from multiplexnanopore_b200 import savemoney as sm
    rng = np.random.default_rng(91)
    refs = [util.rand_seq(rng, int(rng.integers(5000, 7000))) for _ in range(2)]
    fastq = {}
    for k in range(4):
        src = refs[k % 2]
        a = int(rng.integers(len(src)))
        read = util.mutate(rng, src[a:] + src[:a])[: int(rng.integers(5000, 8000))]
        if k == 1:
            read = read.translate(str.maketrans("ACGT", "TGCA"))[::-1]
        fastq[f"long{k}"] = [read, [30] * len(read)]
    fastq["exact_fragment"] = [refs[0][100:700], [30] * 600]          # identical score in both copies of ref + ref
    fastq["gap_first"] = [refs[1][:40] + refs[1][43:400], [30] * 397]  # a deletion right behind the start
    params = dict(gap_open_penalty=3, gap_extend_penalty=1, match_score=1, mismatch_score=-2, topology_of_dna=0)
    stats = {}
    got = sm.legacy_dense_align(refs, fastq, params, engine=engine, stats=stats)
    packed = sum(stats["class_cells"][10:14])
    assert packed > 0.95 * stats["cells"], "the large local problems must run on the packed kernel"
    comp = str.maketrans("ACGT", "TGCA")
    for qid, (read, _) in fastq.items():
        for t, ref in enumerate(refs):
            for strand, q in enumerate((read, read.translate(comp)[::-1])):
                o = orc.align(3, q, ref + ref, 3, 1, 1, -2)
                g = got[qid][2 * t + strand]
                assert (g.score, g.cigar, g.beg_query, g.beg_ref, g.end_query, g.end_ref) == \
                    (o.score, o.cigar_rle, o.beg_query, o.beg_ref, o.end_query, o.end_ref), (qid, t, strand)
    g = got["exact_fragment"][0]
    assert (g.score, g.beg_ref, g.end_ref) == (600, 100, 699)   # first copy
