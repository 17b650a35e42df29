"""Seeded synthetic inputs for the configurations of BASELINE.json.

The reference's own demo FASTQ files are missing blobs (.MISSING_LARGE_BLOBS), so every parity
case and every bench run uses these generators (numpy.random.default_rng(seed), fixed seeds).
"""
from __future__ import annotations

import numpy as np

_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
_COMP = np.zeros(256, dtype=np.uint8)
for _a, _b in zip(b"ACGTMRWSYKVHDBN", b"TGCAKYWSRMBDHVN"):
    _COMP[_a] = _b


def random_bases(rng, n: int) -> np.ndarray:
    return _ACGT[rng.integers(0, 4, size=n)]


def revcomp_bytes(a: np.ndarray) -> np.ndarray:
    return _COMP[a][::-1]


def mutate_bytes(rng, a: np.ndarray, sub=0.03, ins=0.025, dele=0.025) -> np.ndarray:
    """i.i.d. per-base errors: substitution / insertion (after the base) / deletion."""
    n = a.size
    r = rng.random(n)
    out = a.copy()
    is_sub = r < sub
    out[is_sub] = _ACGT[rng.integers(0, 4, size=int(is_sub.sum()))]
    is_ins = (r >= sub) & (r < sub + ins)
    is_del = (r >= sub + ins) & (r < sub + ins + dele)
    reps = np.ones(n, dtype=np.int64)
    reps[is_ins] = 2
    reps[is_del] = 0
    res = np.repeat(out, reps)
    # the second copy of an inserted position becomes a random base
    pos = np.cumsum(reps) - 1
    ins_pos = pos[is_ins]
    res[ins_pos] = _ACGT[rng.integers(0, 4, size=ins_pos.size)]
    if res.size == 0:
        res = a[:1].copy()
    return res


def plasmid_set(rng, n_plasmids=6, len_lo=5000, len_hi=10000, backbone=2500) -> list[str]:
    """C2: plasmids sharing one backbone plus unique inserts (uniform ACGT)."""
    bb = random_bases(rng, backbone)
    out = []
    for _ in range(n_plasmids):
        total = int(rng.integers(len_lo, len_hi + 1))
        ins = random_bases(rng, max(total - backbone, 1))
        out.append(np.concatenate([bb, ins]).tobytes().decode())
    return out


def reads_for(rng, plasmids: list[str], n_reads: int, topology: int = 0, sub=0.03, ins=0.025, dele=0.025,
              frac_full=0.80, frac_partial=0.15, homopolymer_rate=0.0):
    """ONT-like reads (~8 % error): 80 % full length, 15 % fragments (0.3-1.0 N), 5 % concatemers
    (1.2-2.0 N, exercises omit_too_long); random rotation (circular) and strand.
    Returns (reads: list[str], truth: list[(plasmid_idx, is_rc)])."""
    P = [np.frombuffer(p.encode(), dtype=np.uint8) for p in plasmids]
    reads, truth = [], []
    for _ in range(n_reads):
        pi = int(rng.integers(len(P)))
        src = P[pi]
        n = src.size
        u = rng.random()
        if topology == 0:
            k = int(rng.integers(n))
            rot = np.concatenate([src[k:], src[:k]])
            if u < frac_full:
                frag = rot
            elif u < frac_full + frac_partial:
                frag = rot[: max(50, int(n * rng.uniform(0.3, 1.0)))]
            else:
                L = int(n * rng.uniform(1.2, 2.0))
                frag = np.tile(rot, 3)[:L]
        else:
            a = int(rng.integers(0, n // 4 + 1))
            b = int(rng.integers(3 * n // 4, n + 1))
            frag = src[a:b]
        if homopolymer_rate > 0 and frag.size > 40:
            # stretch a few random positions into homopolymers (tie-heavy DP, Appendix E.5)
            frag = frag.copy()
            for _h in range(int(homopolymer_rate * frag.size / 1000) + 1):
                p0 = int(rng.integers(0, frag.size - 12))
                frag[p0:p0 + int(rng.integers(4, 12))] = frag[p0]
        is_rc = bool(rng.random() < 0.5)
        if is_rc:
            frag = revcomp_bytes(frag)
        reads.append(mutate_bytes(rng, frag, sub, ins, dele).tobytes().decode())
        truth.append((pi, is_rc))
    return reads, truth


def config_c2(seed=2, n_reads=20000):
    rng = np.random.default_rng(seed)
    plasmids = plasmid_set(rng)
    reads, truth = reads_for(rng, plasmids, n_reads, topology=0)
    return plasmids, reads, truth


def config_c3(seed=3, n_maps=96, families=8, len_lo=2000, len_hi=15000):
    """pre_survey: families of related maps differing by point edits / indels, random rotation+strand."""
    rng = np.random.default_rng(seed)
    maps = []
    per = n_maps // families
    for _ in range(families):
        base = random_bases(rng, int(rng.integers(len_lo, len_hi + 1)))
        for _v in range(per):
            n_edit = int(rng.integers(1, 51))
            rate = n_edit / base.size
            v = mutate_bytes(rng, base, rate * 0.6, rate * 0.2, rate * 0.2)
            k = int(rng.integers(v.size))
            v = np.concatenate([v[k:], v[:k]])
            if rng.random() < 0.5:
                v = revcomp_bytes(v)
            maps.append(v.tobytes().decode())
    return maps


def config_c4(seed=4, n_plasmids=12, n_reads=200000):
    rng = np.random.default_rng(seed)
    plasmids = plasmid_set(rng, n_plasmids, 3000, 20000, backbone=2000)
    reads, truth = reads_for(rng, plasmids, n_reads, topology=1)
    return plasmids, reads, truth
