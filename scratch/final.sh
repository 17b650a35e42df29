#!/bin/bash
# default bench, then launch list + select capture of the final code (each after its plain run exited 0)
bash scratch/full.sh
export SMX_TRACE_BUDGET_GB=40
CMD="python bench.py --reads 1000 --steps 1 --warmup 1 --batch 1000 --overlap 1 --no-cpu-baseline"
$CMD > gpurun_out/prof_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/prof_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01g.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:smx_select_kernel -s 1 -c 1 -o gpurun_out/prof_select_r01b -f $CMD > gpurun_out/ncu_sel.log 2>&1
echo "select capture rc=$?"
