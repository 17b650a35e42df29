#!/bin/bash
for th in 16 4 1; do
python - $th <<'PY'
import subprocess, sys, json, resource, time
th = sys.argv[1]
t0 = time.time()
p = subprocess.run(["python", "bench.py", "--reads", "8000", "--steps", "4", "--warmup", "1", "--no-cpu-baseline", "--threads", th], capture_output=True, text=True)
wall = time.time() - t0
ru = resource.getrusage(resource.RUSAGE_CHILDREN)
d = json.loads(p.stdout.strip().splitlines()[-1])
print(f"threads {th}: {round(d['ms_per_step'])} ms/step {round(d['e2e']['reads_per_s'])} reads/s; process wall {wall:.1f} s user {ru.ru_utime:.1f} s sys {ru.ru_stime:.1f} s")
PY
done
