#!/bin/bash
# Round-2 ncu evidence: launch list of one step, full captures of the forward kernels (one-warp launch + the team launch
# that runs next to it), traceback, chain.  Every command is run plain first.
export SMX_TRACE_BUDGET_GB=40
CMD="python bench.py --reads 1000 --steps 1 --warmup 1 --batch 1000 --overlap 1 --no-cpu-baseline"
$CMD > gpurun_out/prof_plain_r02.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/prof_plain_r02.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02.csv $CMD > gpurun_out/ncu_launches_r02.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:smx_dp_forward16h -s 2 -c 2 -o gpurun_out/prof_fwd16h_r02 -f $CMD > gpurun_out/ncu_fwd_r02.log 2>&1
echo "forward capture rc=$?"; grep -E "smx_dp_forward16h" gpurun_out/ncu_fwd_r02.log | tail -4
ncu --set full --clock-control none -k regex:smx_traceback -s 1 -c 1 -o gpurun_out/prof_tb_r02 -f $CMD > gpurun_out/ncu_tb_r02.log 2>&1
echo "tb capture rc=$?"
