"""T0 for row A3: CUDA diagonal scan == restated Cython scan (alignment_functions.pyx:46-64)
on the windows ref_query_alignment.py:351-366 builds."""
import numpy as np
import pytest

from oracle import oracle as orc
from tests import util

pytestmark = pytest.mark.gpu


def oracle_counts(ref, q, topology):
    rv = np.array([ord(c) for c in ref], dtype=np.int64)
    qv = np.array([ord(c) for c in q], dtype=np.int64)
    nr, nq = len(ref), len(q)
    if topology == 0:
        repeat = 1 + int(np.ceil((nq - 1) / nr))
        return orc.kmer_offset_scan(np.tile(rv, repeat), qv, nr, nq)
    pad = np.zeros(nq - 1, dtype=np.int64)
    return orc.kmer_offset_scan(np.concatenate([pad, rv, pad]), qv, nr + nq - 1, nq)


@pytest.mark.parametrize("topology", [0, 1])
def test_scan_matches_oracle(engine, topology):
    rng = np.random.default_rng(21 + topology)
    seqs, pairs = [], []
    for _ in range(24):
        nr = int(rng.integers(1, 2500))
        ref = util.rand_seq(rng, nr)
        kind = rng.random()
        if kind < 0.5:
            k = int(rng.integers(0, nr))
            q = util.mutate(rng, ref[k:] + ref[:k])
        elif kind < 0.8:
            q = util.rand_seq(rng, int(rng.integers(1, 4000)))
        else:
            q = util.mutate(rng, ref + ref + ref[: nr // 2])  # longer than the reference (repeat >= 3)
        seqs += [ref, q]
        pairs.append((len(seqs) - 2, len(seqs) - 1))
    offs, lens = engine.upload_sequences(seqs)
    r = np.array([p[0] for p in pairs]); q = np.array([p[1] for p in pairs])
    counts, coff = engine.scan_batch(offs[r], lens[r], offs[q], lens[q], topology)
    for p, (ri, qi) in enumerate(pairs):
        want = oracle_counts(seqs[ri], seqs[qi], topology)
        got = counts[coff[p]: coff[p + 1]]
        assert got.shape == want.shape
        assert np.array_equal(got, want), f"pair {p} nr={len(seqs[ri])} nq={len(seqs[qi])} topology={topology}"


def test_scan_non_acgt_bytes(engine):
    rng = np.random.default_rng(23)
    ref = util.rand_seq(rng, 900, b"ACGTNNRY")
    q = util.mutate(rng, ref[300:] + ref[:300])
    q2 = util.rand_seq(rng, 500, b"ACGTN")
    offs, lens = engine.upload_sequences([ref, q, q2])
    for topology in (0, 1):
        counts, coff = engine.scan_batch(offs[[0, 0]], lens[[0, 0]], offs[[1, 2]], lens[[1, 2]], topology)
        for p, qq in enumerate((q, q2)):
            assert np.array_equal(counts[coff[p]: coff[p + 1]], oracle_counts(ref, qq, topology))
