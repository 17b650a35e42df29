#!/bin/bash
for v in "$@"; do
echo "== $v"
SMX_LIB_PATH=scratch/var/$v.so python scratch/dpbench.py 3000x4000x2400 600x3000x12000 2>&1 | grep -v "W<=4\|W<=8"
done
