#!/bin/bash
export SMX_TRACE_BUDGET_GB=40
CMD="python bench.py --reads 1000 --steps 1 --warmup 1 --batch 1000 --overlap 1 --no-cpu-baseline"
$CMD > gpurun_out/prof_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/prof_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01e.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:smx_dp_forward16h -s 1 -c 1 -o gpurun_out/prof_dp16h_w1_r01 -f $CMD > gpurun_out/ncu_full.log 2>&1
echo "full capture rc=$?"; grep -E "smx_dp_forward16h|PROF" gpurun_out/ncu_full.log | tail -3
ncu --set full --clock-control none -k regex:smx_scan_kernel -s 1 -c 1 -o gpurun_out/prof_scan_r01 -f $CMD > gpurun_out/ncu_scan.log 2>&1
echo "scan capture rc=$?"
ncu --set full --clock-control none -k regex:smx_traceback -s 1 -c 1 -o gpurun_out/prof_tb_r01 -f $CMD > gpurun_out/ncu_tb.log 2>&1
echo "tb capture rc=$?"
