import sys, time, threading, os
sys.path.insert(0, ".")
import numpy as np
from multiplexnanopore_b200 import Engine, synth, savemoney as sm, engine as eng_mod
PARAMS = dict(gap_open_penalty=3, gap_extend_penalty=1, match_score=1, mismatch_score=-2, topology_of_dna=0)
plasmids, reads, _ = synth.config_c2(seed=2, n_reads=20000)
eng = Engine(0)
acc = {}
lock = threading.Lock()
def timed(name, fn):
    def w(*a, **k):
        t = time.perf_counter(); r = fn(*a, **k); dt = time.perf_counter() - t
        with lock: acc[name] = acc.get(name, 0.0) + dt
        return r
    return w
eng_mod._pack = timed("_pack", eng_mod._pack)
sm._merge = timed("_merge", sm._merge)
orig_run = eng_mod.Engine.pipeline_run
lib = eng._lib
class LibProxy:
    def __init__(self, lib): self._l = lib
    def __getattr__(self, n):
        f = getattr(self._l, n)
        if n in ("smx_pipeline_run", "smx_pipeline_fetch", "smx_pipeline_stats"): return timed(n, f)
        return f
for _ in range(2): sm.align_all(plasmids, reads, PARAMS, True, engine=eng)
for e in sm._engine_pools[id(eng)]: e._lib = LibProxy(e._lib)
acc.clear()
t0 = time.perf_counter()
for _ in range(2): res = sm.align_all(plasmids, reads, PARAMS, True, engine=eng)
wall = time.perf_counter() - t0
print("wall per step", wall / 2, {k: round(v / 2, 3) for k, v in acc.items()})
print({k: round(res.stats[k], 3) for k in ("t_total", "t_pool", "t_scan", "t_select", "t_chain", "t_dp", "t_stitch", "gpu_scan_ms", "gpu_select_ms", "gpu_dp_ms")})
