#!/bin/bash
# GPU tests + short bench; prints the serial-pass stage seconds (t_stitch etc.)
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --reads 8000 --steps 2 --warmup 2 --no-cpu-baseline 2>gpurun_out/b2.err | tail -1 > gpurun_out/b2.json
tail -c 300 gpurun_out/b2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/b2.json').read())
print(round(d["ms_per_step"]), "ms/step", round(d["e2e"]["reads_per_s"]), "reads/s; dev GCUPS", round(d["value"]), "e2e GCUPS", round(d["e2e"]["value"]), "util", d["clocks"]["gpu_util_pct_mean"])
print({k:round(v,4) for k,v in d["stage_seconds_per_step_rank0"].items()})
print({k:round(v,1) for k,v in d["kernel_ms_per_step_rank0"].items()})
PY
