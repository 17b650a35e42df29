"""developer tool: forward16 throughput per team width on uniform problem sets"""
import os, sys, time
import numpy as np
sys.path.insert(0, ".")
from multiplexnanopore_b200 import Engine
from tests import util
eng = Engine(0)
rng = np.random.default_rng(1)
shapes = [(3000, 4000, 2400), (600, 3000, 12000), (7600, 8500, 500), (300, 1000, 40000)]
if len(sys.argv) > 1:
    shapes = [tuple(int(x) for x in a.split("x")) for a in sys.argv[1:]]
for (m, n, cnt) in shapes:
    ref = util.rand_seq(rng, n)
    q = util.rand_seq(rng, m)
    pool, s1o, s1l, s2o, s2l = util.pack_problems([(q, ref)])
    eng.upload_pool(pool)
    rep = lambda a: np.repeat(a, cnt)
    for w in (1, 2, 4, 8):
        os.environ["SMX_S16_MAXW"] = str(w)
        eng.align_prepare(rep(s1o), rep(s1l), rep(s2o), rep(s2l), np.ones(cnt, np.uint8), 3, 1, 1, -2)
        best = 1e9
        for it in range(3):
            eng.align_run()
            ms, cells = eng.align_kernel_ms()
            fwd = ms[:14].sum()
            best = min(best, fwd)
        print(f"{m}x{n} x{cnt} W<={w}: forward {best:.2f} ms -> {eng.align_cells/best/1e9:.3f} TCUPS; tb {ms[14]:.2f} ms; classes {[(c, round(ms[c],2)) for c in range(14) if ms[c]>0]}", flush=True)
