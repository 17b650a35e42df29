#!/bin/bash
export SMX_S16_MAXW=2
ncu --set full --clock-control none --import-source on -k regex:smx_dp_forward16h -s 1 -c 1 -o gpurun_out/prof_f16h_r01 -f python scratch/dpbench_one.py > gpurun_out/ncu_f16.log 2>&1
tail -5 gpurun_out/ncu_f16.log
