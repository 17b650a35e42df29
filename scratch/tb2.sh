#!/bin/bash
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python scratch/dpbench.py 3000x4000x2400 600x3000x12000 7600x8500x500 2>&1 | grep -v "W<=8"
python bench.py --reads 8000 --steps 2 --warmup 2 --no-cpu-baseline 2>gpurun_out/b1.err | tail -1 > gpurun_out/b1.json
tail -c 400 gpurun_out/b1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/b1.json').read())
print(round(d["ms_per_step"]), "ms/step", round(d["e2e"]["reads_per_s"]), "reads/s; dev GCUPS", round(d["value"]), "e2e GCUPS", round(d["e2e"]["value"]))
print({k:round(v,3) for k,v in d["stage_seconds_per_step_rank0"].items()})
print({k:round(v,1) for k,v in d["kernel_ms_per_step_rank0"].items()})
print("tcups", d["roofline"]["kernel_tcups"], d["roofline"]["kernel"])
PY
