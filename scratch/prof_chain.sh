#!/bin/bash
export SMX_TRACE_BUDGET_GB=40
CMD="python bench.py --reads 1000 --steps 1 --warmup 1 --batch 1000 --overlap 1 --no-cpu-baseline"
ncu --set full --clock-control none --import-source on -k regex:smx_chain_kernel -s 7 -c 1 -o gpurun_out/prof_chain_r01 -f $CMD > gpurun_out/ncu_chain.log 2>&1
echo "chain capture rc=$?"
