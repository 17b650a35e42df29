#!/bin/bash
# GPU tests, host CPU seconds per stage (one batch in flight), short overlapped bench
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
SMX_PHASE_LOG=1 python bench.py --reads 4000 --steps 1 --warmup 1 --overlap 1 --no-cpu-baseline 2>gpurun_out/phase_o1.log | tail -1 > gpurun_out/b_o1.json
grep "^CPU" gpurun_out/phase_o1.log | awk '{a[$3]+=$4; n[$3]++} END{for(k in a) printf "%s %.1f ms/batch (%d)\n", k, 1000*a[k]/n[k], n[k]}'
python bench.py --reads 8000 --steps 2 --warmup 2 --no-cpu-baseline 2>gpurun_out/b2.err | tail -1 > gpurun_out/b2.json
tail -c 300 gpurun_out/b2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/b2.json').read())
print(round(d["ms_per_step"]), "ms/step", round(d["e2e"]["reads_per_s"]), "reads/s; dev GCUPS", round(d["value"]), "e2e GCUPS", round(d["e2e"]["value"]), "util", d["clocks"]["gpu_util_pct_mean"], "h2d", d["e2e"]["h2d_bytes_per_step"])
print({k:round(v,4) for k,v in d["stage_seconds_per_step_rank0"].items()})
PY
