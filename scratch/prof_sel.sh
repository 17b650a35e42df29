#!/bin/bash
export SMX_TRACE_BUDGET_GB=40
CMD="python bench.py --reads 1000 --steps 1 --warmup 1 --batch 1000 --overlap 1 --no-cpu-baseline"
$CMD > gpurun_out/prof_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/prof_plain.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:smx_select_kernel -s 1 -c 1 -o gpurun_out/prof_select_r01 -f $CMD > gpurun_out/ncu_sel.log 2>&1
echo "select capture rc=$?"
ncu --set full --clock-control none --import-source on -k regex:smx_dp_forward -s 2 -c 1 -o gpurun_out/prof_dpk1_r01 -f $CMD > gpurun_out/ncu_k1.log 2>&1
echo "k1 capture rc=$?"; grep -E "smx_dp_forward" gpurun_out/ncu_k1.log | tail -3
