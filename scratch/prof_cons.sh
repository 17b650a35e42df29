#!/bin/bash
mkdir -p gpurun_out
python scratch/prof_cons.py 2>&1 | tail -2
ncu --set full --clock-control none --import-source on -k regex:smx_consensus_kernel -s 1 -c 1 -o gpurun_out/prof_consensus_r02 -f python scratch/prof_cons.py > gpurun_out/ncu_cons.log 2>&1
echo "capture rc=$?"
