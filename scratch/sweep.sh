#!/bin/bash
for ov in 1 2 3; do for b in 500 1000; do
python bench.py --reads 4000 --steps 2 --warmup 1 --batch $b --overlap $ov --no-cpu-baseline 2>&1 | tail -1 > /tmp/line.json
python - "$ov" "$b" <<'PY'
import json,sys
d=json.loads(open('/tmp/line.json').read())
print("overlap",sys.argv[1],"batch",sys.argv[2], round(d["ms_per_step"]), "ms/step", round(d["e2e"]["reads_per_s"]), "reads/s; e2e GCUPS", round(d["e2e"]["value"]), "dev GCUPS", round(d["value"]))
PY
done; done
