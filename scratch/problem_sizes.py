"""Harvests the DP problems the reference algorithm makes for C2 reads (CPU oracle port, DPLog) and evaluates the
instruction cost model of smx_dp_forward16h on them: band padding, share of pairs without a high band, share of the
general-loop chunks, single-warp time of the largest problems.  Developer tool behind DESIGN.md 3.1 / 7; CPU only.

    python scratch/problem_sizes.py [n_reads]
"""
import sys

import numpy as np

sys.path.insert(0, ".")
from multiplexnanopore_b200 import synth  # noqa: E402
from oracle import oracle as orc  # noqa: E402
from oracle import savemoney_port as port  # noqa: E402

PARAMS = dict(gap_open_penalty=3, gap_extend_penalty=1, match_score=1, mismatch_score=-2, topology_of_dna=0)
STEP8, STEP4, GENERAL = 152.0, 85.0, 470.0  # SASS instructions per pair step: K = 8 interior, K = 4 interior, general loop
LAG = 64


def main():
    n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 36
    orc.build()
    plasmids, reads, _ = synth.config_c2(seed=2, n_reads=n_reads)
    calls = []
    for r in reads:
        log = port.DPLog()
        port.align_read(plasmids, r, PARAMS, True, log)
        calls += log.calls
    a = np.array(calls)
    tiny, big = a[a[:, 1] <= 32], a[a[:, 1] > 32]
    rows, cols = big[:, 1].astype(float), big[:, 2].astype(float)
    real = (rows * cols).sum()
    print(f"{len(a)} problems of {n_reads} reads: {len(tiny)} of <= 32 rows ({(tiny[:, 1] * tiny[:, 2]).sum() / 1e9:.4f} G cells, "
          f"longest reference slice {tiny[:, 2].max()}), {len(big)} packed-class ({real / 1e9:.2f} G cells)")
    print("rows of the packed class: quantiles 10/50/90/99 =", np.percentile(rows, [10, 50, 90, 99]).astype(int),
          " cell-weighted mean", int((rows * rows * cols).sum() / real))
    nb = np.ceil(rows / 256)
    npairs = np.ceil(nb / 2)
    steps = cols + 31 + LAG
    lone = (nb % 2 == 1)
    print(f"computed / real cells at 512-row pairs: {(npairs * 512 * cols).sum() / real:.3f}")
    print(f"pairs without a high band: {(lone * steps).sum() / (npairs * steps).sum():.3f} of all pair steps")
    k8 = (npairs * steps * STEP8).sum()
    rem = rows - (nb - 1) * 256
    k4 = ((npairs - lone) * steps * STEP8 + lone * np.where(rem > 128, steps, cols + 31) * STEP4).sum()
    ideal = (rows / 512 * steps * STEP8).sum()
    print(f"instruction cost: lone bands at K = 4 {k4 / k8:.3f}, padding-free {ideal / k8:.3f} (K = 8 everywhere = 1)")
    chunks = np.ceil(steps / 32)
    gen = np.minimum(chunks, np.where(nb == 1, 3, 7))  # entry + exit chunks: 3 + 4 with a high band, 1 + 2 for a single band (rough)
    share = (npairs * gen).sum() / (npairs * chunks).sum()
    print(f"general-loop chunks: {share:.3f} of the pair steps, {share * GENERAL / (share * GENERAL + (1 - share) * STEP8):.3f} of the instructions")
    order = np.argsort(-(rows * cols))[:5]
    for i in order:  # one warp, scheduler shared by four warps: ~930 cycles per pair step at 1.965 GHz
        ms = npairs[i] * steps[i] * 930 / 1.965e9 * 1e3
        print(f"  {int(rows[i])} x {int(cols[i])}: {npairs[i] * steps[i] / 1e3:.0f} k pair steps ~ {ms:.0f} ms on its warp")


if __name__ == "__main__":
    main()
