#!/bin/bash
# two GPUs: multi-GPU parity tests, then strong and weak scaling of C2 and the in-process devices= path
mkdir -p gpurun_out
python -m pytest tests/test_multi_gpu.py tests/test_outputs_gpu.py -m gpu -x -q > gpurun_out/n2_pytest.log 2>&1; tail -4 gpurun_out/n2_pytest.log
for mode in strong weak; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 2 --scaling $mode > gpurun_out/n2_$mode.json 2> gpurun_out/n2_$mode.err
  echo "rc=$?"
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/n2_$mode.json")); print("N=2 $mode: value %.0f e2e %.0f GCUPS %.0f reads/s ms/step %.0f" % (d["value"], d["e2e"]["value"], d["e2e"]["reads_per_s"], d["ms_per_step"]))
except Exception as ex: print("N=2 $mode failed", ex); print(open("gpurun_out/n2_$mode.err").read()[-2500:])
PY
done
python - <<'PY'
# the product path: one process, devices=[0, 1] (threads), 20 000 C2 reads
import time, numpy as np
from multiplexnanopore_b200 import savemoney as sm, synth
plasmids, reads, _ = synth.config_c2(seed=2, n_reads=20000)
P = dict(gap_open_penalty=3, gap_extend_penalty=1, match_score=1, mismatch_score=-2, topology_of_dna=0)
for devs in ([0], [0, 1]):
    sm.align_all(plasmids, reads[:4000], P, True, devices=devs)
    t = time.perf_counter(); r = sm.align_all(plasmids, reads, P, True, devices=devs); dt = time.perf_counter() - t
    print(f"in-process devices={devs}: {len(reads) / dt:.0f} reads/s, {r.stats['dp_cells'] / dt / 1e9:.0f} GCUPS")
PY
