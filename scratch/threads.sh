#!/bin/bash
for th in 16 8 4 2; do
python bench.py --reads 4000 --steps 2 --warmup 1 --no-cpu-baseline --threads $th 2>&1 | tail -1 > /tmp/line.json
python - "$th" <<'PY'
import json,sys
d=json.loads(open('/tmp/line.json').read())
print("threads",sys.argv[1], round(d["ms_per_step"]), "ms/step", round(d["e2e"]["reads_per_s"]), "reads/s;", {k:round(v,3) for k,v in d["stage_seconds_per_step_rank0"].items()})
PY
done
