#!/bin/bash
for cfg in "-DSMX_TB_PF=0" "-DSMX_TB_PF=1 -DSMX_TB_AHEAD=48" "-DSMX_TB_PF=2 -DSMX_TB_AHEAD=64" "-DSMX_TB_PF=3 -DSMX_TB_AHEAD=16" "-DSMX_TB_PF=2 -DSMX_TB_AHEAD=256"; do
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -shared -Xcompiler -fPIC -cudart static $cfg -o multiplexnanopore_b200/libsmx.so multiplexnanopore_b200/csrc/smx_api.cu
echo "$cfg"; bash scratch/stages.sh | sed -n '3p' | grep -o "'traceback': [0-9.]*"
done
