#!/bin/bash
# full capture of ALL forward launches of the timed step (team launches + the one-warp launch)
export SMX_TRACE_BUDGET_GB=40
CMD="python bench.py --reads 1000 --steps 1 --warmup 1 --batch 1000 --overlap 1 --no-cpu-baseline"
$CMD > gpurun_out/prof_plain_r02b.log 2>&1 || { echo "plain run failed"; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:smx_dp_forward16h -s 3 -c 3 -o gpurun_out/prof_fwd16h_r02b -f $CMD > gpurun_out/ncu_fwd_r02b.log 2>&1
echo "forward capture rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02b.csv $CMD > /dev/null 2>&1
echo "launch list rc=$?"
