#!/bin/bash
# ncu evidence for the checkpointed path: launch list of one step, full captures of the score-only forward launch (one warp
# per problem) and of smx_traceback_ck.  Every command is run plain first.
export SMX_TRACE_BUDGET_GB=40
CMD="python bench.py --reads 1000 --steps 1 --warmup 1 --batch 1000 --overlap 1 --no-cpu-baseline"
$CMD > gpurun_out/prof_plain_ck.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/prof_plain_ck.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02_ck.csv $CMD > gpurun_out/ncu_launches_ck.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:forward16hILi1ELb0ELb1 -s 1 -c 1 -o gpurun_out/prof_fwd16h_ck -f $CMD > gpurun_out/ncu_fwd_ck.log 2>&1
echo "forward capture rc=$?"
ncu --set full --clock-control none --import-source on -k regex:smx_traceback_ck -s 1 -c 1 -o gpurun_out/prof_tb_ck -f $CMD > gpurun_out/ncu_tb_ck.log 2>&1
echo "tb capture rc=$?"
ls -la gpurun_out/*.ncu-rep
