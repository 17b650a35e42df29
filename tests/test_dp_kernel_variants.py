"""The three forward-DP kernel families must agree with the oracle (and therefore with each other) on the same
problems: smx_dp_forward16h (u16x2 diagonal frame, default), smx_dp_forward16 (s16x2, SMX_S16H=0 or scorings the
diagonal frame cannot take) and smx_dp_forward (int32, SMX_S16=0).  Plus the ends of the packed kernel's value range
(long, highly similar and completely dissimilar pairs) and its launch modes."""
import os

import numpy as np
import pytest

from oracle import oracle as orc
from tests import util
from tests.test_dp_parity import run_and_compare

pytestmark = pytest.mark.gpu


class _Env:
    def __init__(self, **kv):
        self.kv, self.old = kv, {}

    def __enter__(self):
        for k, v in self.kv.items():
            self.old[k] = os.environ.get(k)
            os.environ[k] = v

    def __exit__(self, *exc):
        for k, v in self.old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def _mixed_problems(seed, n=40, max_len=1400):
    rng = np.random.default_rng(seed)
    probs = util.make_problems(rng, n, max_len, related=0.75)
    modes = rng.integers(0, 3, size=len(probs))
    return probs, modes


@pytest.mark.parametrize("env", [dict(SMX_S16H="0"), dict(SMX_S16="0"), dict(SMX_PERSISTENT="1"), dict(SMX_S16_MINROWS="256"),
                                 dict(SMX_S16_MAXW="1"), dict(SMX_S16_MAXW="8"), dict(SMX_TEAM_MIN_MCELLS="1"), dict(SMX_S16H_SPLIT="0")])
def test_kernel_families_and_launch_modes(engine, env):
    probs, modes = _mixed_problems(61)
    with _Env(**env):
        run_and_compare(engine, probs, modes, (3, 1, 1, -2))
        run_and_compare(engine, probs[:12], modes[:12], (5, 2, 2, -3))


@pytest.mark.parametrize("scoring", [(1, 3, 1, -2),      # go < ge: the constant of the diagonal frame changes sign
                                     (3, 1, 1, -5),      # mismatch + 2 ge < 0: falls back to the s16x2 kernel
                                     (40, 20, 60, -30),  # match + go + ge = 120: profile bytes near their limit
                                     (60, 30, 40, -50),  # 130 > 127: s16x2 kernel
                                     (2, 1, 0, -1)])     # match 0
def test_scorings_around_the_packed_kernel_conditions(engine, scoring):
    probs, modes = _mixed_problems(62, n=24, max_len=900)
    run_and_compare(engine, probs, modes, scoring)


def test_value_range_ends(engine):
    """12 000 x 13 000: scores near +12 000 (similar pair) and near -(go + 25 000 ge) along the borders (dissimilar pair);
    with (5, 2, 2, -3) the biased diagonal-frame values pass 65 000 and the problem must leave the packed kernel."""
    rng = np.random.default_rng(63)
    ref = util.rand_seq(rng, 13000)
    similar = util.mutate(rng, ref, 0.01, 0.005, 0.005)[:12000]
    other = util.rand_seq(rng, 12000)
    low = "AC" * 6000  # tie-heavy against itself shifted
    probs = [(similar, ref), (other, ref), (similar, ref), (other, ref), (low, low[1:] + "A")]
    modes = [0, 0, 1, 2, 0]
    run_and_compare(engine, probs, modes, (3, 1, 1, -2))
    run_and_compare(engine, probs[:2], modes[:2], (5, 2, 2, -3))


@pytest.mark.parametrize("split", ["1", "0"])
def test_split_last_band_row_counts(engine, split):
    """A last band without a partner runs as two half-bands of 128 rows at 4 rows per lane (smx_dp16h.cuh) and keeps its
    direction bits in the K = 4 layout, which the traceback enters and leaves: row counts around every boundary of
    that scheme (one half-band, two half-bands, one and two full pairs in front), short and long references, all three
    global / semi-global modes, with the split and without it (SMX_S16H_SPLIT=0)."""
    rng = np.random.default_rng(63)
    probs, modes = [], []
    for m in (33, 127, 128, 129, 130, 255, 256, 257, 511, 512, 513, 639, 640, 641, 767, 768, 769, 1025, 1153, 1281):
        for n in (40, 97, 700):
            ref = util.rand_seq(rng, max(n, m) + 40)
            q = util.mutate(rng, ref, 0.05, 0.04, 0.04)
            q = (q * (m // max(1, len(q)) + 1))[:m]
            for mode in (0, 1, 2):
                probs.append((q, ref[:n]))
                modes.append(mode)
    with _Env(SMX_S16H_SPLIT=split):
        run_and_compare(engine, probs, np.array(modes), (3, 1, 1, -2))


@pytest.mark.parametrize("env", [dict(), dict(SMX_TEAMS_CONCURRENT="0"), dict(SMX_S16_MAXW="2"), dict(SMX_S16_MAXW="4")])
def test_teams_next_to_the_one_warp_launch(engine, env):
    """The largest problems of a chunk run as teams of 2 / 4 / 8 warps on their own streams while the one-warp-per-problem
    launch takes the rest (fork / join by events); with the thresholds lowered every team width gets problems here:
    2 warps from 4 band pairs, 4 from 8, 8 from 16 (rows 1 600 .. 8 200), semi-global ends included."""
    rng = np.random.default_rng(71)
    probs, modes = [], []
    for m, n in ((1600, 700), (2200, 500), (2300, 900), (3700, 420), (4200, 380), (4100, 610), (7800, 300), (8200, 333), (600, 900), (90, 70), (300, 2000)):
        ref = util.rand_seq(rng, n)
        q = util.mutate(rng, (ref * (m // n + 2))[:m], 0.08, 0.03, 0.03)[:m]
        for mode in (0, 1, 2):
            probs.append((q, ref)); modes.append(mode)
    probs += util.make_problems(rng, 30, 500)
    modes += [int(x) for x in rng.integers(0, 3, size=30)]
    with _Env(SMX_TEAM2_MCELLS="0.5", SMX_TEAM4_MCELLS="1.2", SMX_TEAM8_MCELLS="2.0", **env):
        run_and_compare(engine, probs, modes, (3, 1, 1, -2))


def test_packed_local_kernel_sizes_and_teams(engine):
    """mode 3 (sw_trace) on the packed kernel: every row-count class incl. split last bands, problems with no positive
    cell, repeats (several cells share the best score: smallest reference position, then smallest query position wins),
    wildcards in the query, and the team widths."""
    rng = np.random.default_rng(83)
    probs = []
    for m, n in ((33, 40), (64, 64), (65, 300), (128, 77), (129, 500), (256, 256), (257, 90), (300, 1000), (511, 200), (513, 700), (700, 1300),
                 (1025, 333), (1600, 700), (2300, 900), (4100, 610), (7800, 300)):
        ref = util.rand_seq(rng, n)
        q = util.mutate(rng, (ref * (m // n + 2))[:m], 0.06, 0.03, 0.03)[:m]
        probs.append((q, ref))                       # repeats of the reference inside the query
        probs.append((util.rand_seq(rng, m), ref))   # unrelated
    probs.append(("A" * 200, "C" * 300))             # no positive cell: score 0, empty CIGAR
    probs.append(("ACGT" * 60, "ACGT" * 90))         # many equal maxima
    probs.append((util.rand_seq(rng, 150, b"ACGTN"), util.rand_seq(rng, 400)))
    modes = [3] * len(probs)
    run_and_compare(engine, probs, modes, (3, 1, 1, -2))
    run_and_compare(engine, probs[:16], modes[:16], (5, 2, 2, -3))
    with _Env(SMX_TEAM2_MCELLS="0.5", SMX_TEAM4_MCELLS="1.2", SMX_TEAM8_MCELLS="2.0"):
        run_and_compare(engine, probs, modes, (3, 1, 1, -2))
    with _Env(SMX_S16="0"):   # the int32 local kernels follow the same end-cell order
        run_and_compare(engine, probs[-12:], modes[-12:], (3, 1, 1, -2))


def test_guard_bands_around_direction_bits(engine):
    """compute-sanitizer is closed on this pool (profiles/sanitizer_closed_r02.log); SMX_GUARD=1 surrounds every problem's
    direction-bit region with poisoned bytes and verifies them after the traceback: no forward kernel family, team width,
    split last band or local problem writes outside its region, and the results are still the oracle's."""
    rng = np.random.default_rng(97)
    probs, modes = _mixed_problems(62, n=60, max_len=900)
    probs = list(probs); modes = [int(x) for x in modes]
    for m, n in ((1600, 700), (2300, 500), (4100, 300), (7800, 200), (255, 257), (257, 255), (129, 64), (513, 31)):
        ref = util.rand_seq(rng, n)
        q = util.mutate(rng, (ref * (m // n + 2))[:m], 0.08, 0.03, 0.03)[:m]
        for mode in (0, 1, 2, 3):
            probs.append((q, ref)); modes.append(mode)
    for env in (dict(), dict(SMX_S16H="0"), dict(SMX_S16="0"), dict(SMX_S16H_SPLIT="0"), dict(SMX_TEAM2_MCELLS="0.5", SMX_TEAM4_MCELLS="1.2", SMX_TEAM8_MCELLS="2.0")):
        with _Env(SMX_GUARD="1", **env):
            run_and_compare(engine, probs, modes, (3, 1, 1, -2))
            assert engine.guard_errors == 0, env
    with _Env(SMX_GUARD="1"):
        engine.set_trace_budget(1 << 20)
        try:
            run_and_compare(engine, probs[:80], modes[:80], (3, 1, 1, -2))
        finally:
            engine.set_trace_budget(16 << 30)
        assert engine.guard_errors == 0


def _ck_problems(rng, shapes, modes=(0, 1, 2), low=False):
    probs, ms = [], []
    for m, n in shapes:
        if low:
            unit = util.rand_seq(rng, int(rng.integers(1, 4)))
            ref = (unit * (n // len(unit) + 1))[:n]
        else:
            ref = util.rand_seq(rng, n)
        q = util.mutate(rng, (ref * (m // n + 2))[:m + 40], 0.06, 0.04, 0.04)
        q = (q * (m // max(1, len(q)) + 1))[:m]
        for mode in modes:
            probs.append((q, ref)); ms.append(mode)
    return probs, ms


@pytest.mark.parametrize("env", [dict(), dict(SMX_S16H_SPLIT="0"), dict(SMX_TEAM2_MCELLS="0.5", SMX_TEAM4_MCELLS="1.2", SMX_TEAM8_MCELLS="2.0"),
                                 dict(SMX_GUARD="1")])
def test_checkpointed_traceback_shapes(engine, env):
    """Large packed problems keep checkpoints instead of direction bits (SMX_MODE_CKPT, smx_traceback_ck.cuh): with the
    thresholds lowered to 513 rows x 1 column every shape goes through that path -- odd and even band counts (split last
    bands with one and two half-bands), references shorter than a chunk, than the 64-step lag and than a block, team
    widths, tie-heavy low-complexity pairs -- and must give the oracle's score, end cell and CIGAR."""
    rng = np.random.default_rng(131)
    shapes = [(m, n) for m in (513, 640, 641, 767, 768, 769, 1024, 1025, 1153, 1281, 1537, 2049, 2300) for n in (1, 33, 64, 65, 97, 130, 700, 2100)]
    probs, modes = _ck_problems(rng, shapes)
    p2, m2 = _ck_problems(rng, [(700, 900), (1300, 1100), (1700, 300), (2600, 2800)], low=True)
    probs += p2; modes += m2
    probs += [(util.rand_seq(rng, 1400), util.rand_seq(rng, 1700)), (util.rand_seq(rng, 900), util.rand_seq(rng, 60))]  # unrelated
    modes += [0, 1]
    with _Env(SMX_CKPT_MINROWS="513", SMX_CKPT_MINCOLS="1", **env):
        run_and_compare(engine, probs, modes, (3, 1, 1, -2))
        run_and_compare(engine, probs[::7], modes[::7], (5, 2, 2, -3))
        if "SMX_GUARD" in env:
            assert engine.guard_errors == 0


def test_checkpointed_traceback_default_thresholds(engine):
    """At the default thresholds (>= 513 rows, >= 1 500 columns) next to problems that keep their direction bits; the
    same problems with SMX_CKPT=0 give the same answer (both are checked against the oracle)."""
    rng = np.random.default_rng(137)
    shapes = [(513, 1500), (512, 1600), (1100, 1499), (1024, 3000), (2500, 4000), (3100, 2900), (5200, 6100), (600, 2500), (300, 300), (4000, 1501)]
    probs, modes = _ck_problems(rng, shapes)
    probs += util.make_problems(rng, 40, 800)
    modes += [int(x) for x in rng.integers(0, 4, size=40)]
    run_and_compare(engine, probs, modes, (3, 1, 1, -2))
    with _Env(SMX_CKPT="0"):
        run_and_compare(engine, probs, modes, (3, 1, 1, -2))
    engine.set_trace_budget(8 << 20)   # several chunks, checkpointed and plain problems in each
    try:
        run_and_compare(engine, probs, modes, (3, 1, 1, -2))
    finally:
        engine.set_trace_budget(16 << 30)
