#!/bin/bash
# one GPU's share of a strong-scaled call: 2 500 / 5 000 reads per call, batch size rules
for reads in 2500 5000; do
for cfg in "0 1000" "1 500" "1 334" "1 250" "0 500"; do
  set -- $cfg
  SMX_ADAPT_BATCH=$1 SMX_ADAPT_MIN=$2 python bench.py --reads $reads --steps 4 --warmup 3 --batch $( [ $1 == 0 ] && echo $2 || echo 1000 ) --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith(chr(123))][-1])
print('reads $reads adapt $1 min/batch $2: %.0f reads/s ms/step %.0f' % (d['e2e']['reads_per_s'], d['ms_per_step']))"
done; done
