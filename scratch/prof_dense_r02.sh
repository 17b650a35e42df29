#!/bin/bash
# ncu evidence for the packed local kernel (legacy dense mode): launch list + full capture of the widest forward launch
CMD="python bench.py --config legacy_dense --reads 128 --steps 1 --warmup 1 --no-cpu-baseline"
$CMD > gpurun_out/prof_dense_plain_r02.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/prof_dense_plain_r02.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_dense_r02.csv $CMD > gpurun_out/ncu_launches_dense_r02.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:smx_dp_forward16h -s 2 -c 2 -o gpurun_out/prof_fwd16h_local_r02 -f $CMD > gpurun_out/ncu_fwd_local_r02.log 2>&1
echo "capture rc=$?"; grep -E "smx_dp_forward16h" gpurun_out/ncu_fwd_local_r02.log | tail -3
