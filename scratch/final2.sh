#!/bin/bash
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
bash scratch/full.sh
python bench.py --impl reference --steps 1 --warmup 1 2>/dev/null | tail -1 > gpurun_out/bench_ref.json
python -c "
import json; d=json.loads(open('gpurun_out/bench_ref.json').read()); print('reference arm', round(d['value'],3), d['unit'], round(d['reads_per_s'],1), 'reads/s', d['cpu_baseline']['cores'], 'cores')"
