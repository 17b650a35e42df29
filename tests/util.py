"""Shared helpers of the parity tests: seeded synthetic problems."""
import numpy as np

ALPHA = np.frombuffer(b"ACGT", dtype=np.uint8)


def rand_seq(rng, n, alphabet=b"ACGT"):
    a = np.frombuffer(alphabet, dtype=np.uint8)
    return a[rng.integers(0, len(a), size=n)].tobytes().decode()


def mutate(rng, s, sub=0.03, ins=0.025, dele=0.025):
    out = []
    for c in s:
        r = rng.random()
        if r < sub:
            out.append("ACGT"[rng.integers(4)])
        elif r < sub + ins:
            out.append(c)
            out.append("ACGT"[rng.integers(4)])
        elif r < sub + ins + dele:
            continue
        else:
            out.append(c)
    return "".join(out) or "A"


def make_problems(rng, n, max_len, related=0.7, alphabet=b"ACGT", low_complexity=0.2):
    """(s1=query, s2=ref) pairs: related pairs are noisy copies (realistic), others random;
    a share of low-complexity / homopolymer sequences provokes DP ties."""
    probs = []
    for _ in range(n):
        L = int(rng.integers(1, max_len + 1))
        if rng.random() < low_complexity:
            unit = rand_seq(rng, int(rng.integers(1, 4)))
            ref = (unit * (L // len(unit) + 1))[:L]
        else:
            ref = rand_seq(rng, L, alphabet)
        if rng.random() < related:
            q = mutate(rng, ref, 0.05, 0.04, 0.04)
            if rng.random() < 0.3:  # truncated / extended ends
                a = int(rng.integers(0, max(1, len(q) // 3)))
                q = q[a:] or "C"
        else:
            q = rand_seq(rng, int(rng.integers(1, max_len + 1)), alphabet)
        probs.append((q, ref))
    return probs


def pack_problems(probs):
    """-> (pool bytes, s1_off, s1_len, s2_off, s2_len)"""
    chunks, s1o, s1l, s2o, s2l = [], [], [], [], []
    pos = 0
    for q, r in probs:
        s1o.append(pos); s1l.append(len(q)); chunks.append(q); pos += len(q)
        s2o.append(pos); s2l.append(len(r)); chunks.append(r); pos += len(r)
    return ("".join(chunks).encode(), np.array(s1o, np.int64), np.array(s1l, np.int32),
            np.array(s2o, np.int64), np.array(s2l, np.int32))
