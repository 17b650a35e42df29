#!/bin/bash
# GPU tests, then the host CPU seconds per stage with the run-length and the column-string merge
( time timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 ) 2>&1 | tail -8
for mode in runs chars; do
  SMX_MERGE=$mode SMX_PHASE_LOG=1 python bench.py --reads 8000 --steps 2 --warmup 2 --no-cpu-baseline 2>gpurun_out/phase_$mode.log | tail -1 > gpurun_out/b_$mode.json
  python - $mode <<'PY'
import json, sys, collections
m = sys.argv[1]
d = json.loads(open(f'gpurun_out/b_{m}.json').read())
cpu = collections.defaultdict(float)
for l in open(f'gpurun_out/phase_{m}.log'):
    if l.startswith("CPU "):
        _, h, lab, s = l.split(); cpu[lab] += float(s)
print(m, round(d["ms_per_step"]), "ms/step", round(d["e2e"]["reads_per_s"]), "reads/s", "cpu s by stage (4 passes x 8000 reads + serial pass):", {k: round(v, 2) for k, v in cpu.items()})
PY
done
