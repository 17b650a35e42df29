#!/bin/bash
export SMX_TRACE_BUDGET_GB=40
CMD="python bench.py --reads 1000 --steps 1 --warmup 1 --batch 1000 --overlap 1 --no-cpu-baseline"
$CMD > gpurun_out/prof_plain.log 2>&1 || { echo "plain run failed"; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01f.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
