"""ad-hoc: int peak + DP throughput on a few large problems (developer tool, not the bench)."""
import sys, time, json
import numpy as np
sys.path.insert(0, ".")
from multiplexnanopore_b200 import Engine
from tests import util
eng = Engine(0)
print("int peak", eng.measure_int_peak(4000))
rng = np.random.default_rng(1)
for (m, n, cnt) in [(4000, 8000, 600), (500, 1350, 20000), (31, 31, 400000), (129, 130, 100000)]:
    ref = util.rand_seq(rng, n)
    q = util.mutate(rng, ref)[:m]
    q = (q * (m // len(q) + 1))[:m]
    probs = [(q, ref)]
    pool, s1o, s1l, s2o, s2l = util.pack_problems(probs)
    eng.upload_pool(pool)
    rep = lambda a: np.repeat(a, cnt)
    eng.align_prepare(rep(s1o), rep(s1l), rep(s2o), rep(s2l), np.ones(cnt, np.uint8), 3, 1, 1, -2)
    for it in range(3):
        t = time.time(); eng.align_run(); dt = time.time() - t
        print(f"{m}x{n} x{cnt}: kernel {eng.last_run_ms:.2f} ms wall {dt*1e3:.1f} ms -> {eng.align_cells/eng.last_run_ms/1e6:.1f} GCUPS launches {eng.last_run_launches}")
# scan
seqs = [util.rand_seq(rng, 7500), util.rand_seq(rng, 7500)]
offs, lens = eng.upload_sequences(seqs)
npair = 2000
eng.scan_prepare(np.repeat(offs[:1], npair), np.repeat(lens[:1], npair), np.repeat(offs[1:], npair), np.repeat(lens[1:], npair), 0)
for it in range(3):
    eng.scan_run(); print(f"scan {npair} pairs 7.5k x 7.5k: {eng.last_run_ms:.2f} ms -> {eng.scan_cells/eng.last_run_ms/1e6:.1f} Gcells/s")
