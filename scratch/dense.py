import sys, time
sys.path.insert(0, ".")
from multiplexnanopore_b200 import Engine, synth, savemoney as sm
PARAMS = dict(gap_open_penalty=3, gap_extend_penalty=1, match_score=1, mismatch_score=-2, topology_of_dna=0)
plasmids, reads, _ = synth.config_c2(seed=2, n_reads=64)
fq = {f"r{i}": [r, None] for i, r in enumerate(reads)}
eng = Engine(0)
sm.legacy_dense_align(plasmids, dict(list(fq.items())[:8]), PARAMS, engine=eng)
t = time.perf_counter(); out = sm.legacy_dense_align(plasmids, fq, PARAMS, engine=eng); dt = time.perf_counter() - t
cells = sum(2 * len(r) * sum(2 * len(p) for p in plasmids) for r in reads)
print(f"legacy dense mode: {len(reads)} reads x 6 plasmids x 2 strands, {cells/1e9:.1f} G cells in {dt:.2f} s = {cells/dt/1e9:.0f} GCUPS e2e ({len(reads)/dt:.1f} reads/s)")
