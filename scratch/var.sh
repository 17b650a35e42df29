#!/bin/bash
for v in "$@"; do
echo "== $v"
SMX_LIB_PATH=scratch/var/$v.so python scratch/dpbench.py 3000x4000x4736 2000x3000x9472 2>&1 | grep -v "W<=8" | sed 's/; tb.*//'
done
