#!/bin/bash
for mw in 16 20 24; do
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -shared -Xcompiler -fPIC -cudart static -DSMX16_MIN_WARPS_PER_SM=$mw -o multiplexnanopore_b200/libsmx.so multiplexnanopore_b200/csrc/smx_api.cu
echo "min warps/SM $mw"; bash scratch/stages.sh | sed -n '1p;3p'
done
