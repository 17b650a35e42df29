#!/bin/bash
python bench.py --reads 4000 --steps 2 --warmup 1 --batch 1000 --overlap 1 --no-cpu-baseline 2>&1 | tail -1 > /tmp/line.json
python - <<'PY'
import json
d=json.loads(open('/tmp/line.json').read())
print(round(d["ms_per_step"]), "ms/step", round(d["e2e"]["reads_per_s"]), "reads/s", "dev GCUPS", round(d["value"]))
print({k:round(v,3) for k,v in d["stage_seconds_per_step_rank0"].items()})
print({k:round(v,1) for k,v in d["kernel_ms_per_step_rank0"].items()})
print(d["e2e"]["d2h_bytes_per_step"], d["roofline"]["frac"], d["roofline"]["kernel_tcups"])
PY
