"""Row A4, conserved runs (WindowAnalysis.exec, ref_query_alignment.py:966-984: maximal runs of matches of length > 20
on a selected diagonal).  smx_select_kernel finds them 32 bases at a time: one ballot word per lane, the run entering a
word resolved through the nearest lower word that has a zero bit, the one run of more than 20 matches that fits inside
32 bits by AND-shifts.  This file restates that per-lane logic in Python, instruction for instruction, and checks it
against the plain definition on random and adversarial match strings (the CUDA kernel itself is compared with the host
re-decision and the goldens in the -m gpu tests).  No GPU."""
import numpy as np

M32 = 0xFFFFFFFF
RUN_MIN = 20


def ffs(x):  # CUDA __ffs: 1-based position of the lowest set bit, 0 for x == 0
    return (x & -x).bit_length()


def clz(x):  # CUDA __clz on 32 bits
    return 32 - x.bit_length()


def runs_by_definition(bits):
    out, n, i = [], len(bits), 0
    while i < n:
        if bits[i]:
            j = i
            while j < n and bits[j]:
                j += 1
            if j - i > RUN_MIN:
                out.append((i, j - 1))
            i = j
        else:
            i += 1
    return sorted(out)


def runs_like_the_kernel(bits):
    nq, out = len(bits), []

    def report(rs, re):
        if re - rs + 1 > RUN_MIN:
            out.append((rs, re))

    carry = -1
    for g in range(0, nq, 1024):
        nwords = min(32, (nq - g + 31) >> 5)
        mine = [0] * 32
        for w in range(nwords):
            for lane in range(32):
                i = g + w * 32 + lane
                if i < nq and bits[i]:
                    mine[w] |= 1 << lane
        allones = sum(1 << lane for lane in range(32) if mine[lane] == M32)
        open_out = []
        for lane in range(32):
            nm = ~mine[lane] & M32
            h = clz(nm) if nm else 32
            open_out.append(-1 if h == 0 else g + lane * 32 + 32 - h)
        out31 = -1
        for lane in range(32):
            base = g + lane * 32
            nm = ~mine[lane] & M32
            c = ffs(nm) - 1 if nm else 32
            lower = ~allones & ((1 << lane) - 1)
            src = 31 - clz(lower) if lower else -1
            entering = open_out[src] if src >= 0 else carry
            if entering < 0 and lane - 1 > src:
                entering = g + (src + 1) * 32
            rest = mine[lane]
            if entering >= 0:
                if c < 32:
                    report(entering, base + c - 1)
                rest = 0 if c == 32 else mine[lane] & ~((1 << c) - 1) & M32
            a2 = rest & (rest >> 1)
            a4 = a2 & (a2 >> 2)
            a8 = a4 & (a4 >> 4)
            a16 = a8 & (a8 >> 8)
            a21 = a16 & (a4 >> 16) & (rest >> 20)
            if a21:
                p = ffs(a21) - 1
                inv = ~(rest >> p) & M32
                e = p + ffs(inv) - 2 if inv else 31
                if e < 31:
                    report(base + p, base + e)
            if lane == 31:
                out31 = (entering if entering >= 0 else base) if nm == 0 else open_out[lane]
        carry = out31
    if carry >= 0:
        report(carry, nq - 1)
    return sorted(out)


def test_kernel_logic_equals_definition():
    rng = np.random.default_rng(0)
    for _ in range(1200):
        nq = int(rng.choice([1, 5, 20, 21, 22, 31, 32, 33, 63, 64, 100, 1023, 1024, 1025, 2047, 2048, 3000, int(rng.integers(1, 5000))]))
        density = float(rng.choice([0.25, 0.5, 0.9, 0.97, 0.995, 1.0]))
        bits = (rng.random(nq) < density).astype(int)
        if rng.random() < 0.3:  # a long run somewhere, possibly up to either end
            a = int(rng.integers(0, nq))
            b = int(rng.integers(a, nq + 1))
            bits[a:b] = 1
        assert runs_like_the_kernel(bits) == runs_by_definition(bits), (nq, density)


def test_boundary_lengths():
    for nq in (21, 32, 53, 64, 1024, 1045, 2048):
        for length in (20, 21, 22):
            for start in (0, 1, 11, 31, 32, nq - length):
                if start < 0 or start + length > nq:
                    continue
                bits = np.zeros(nq, int)
                bits[start:start + length] = 1
                assert runs_like_the_kernel(bits) == runs_by_definition(bits), (nq, length, start)
