#!/bin/bash
# 2 GPUs: the multi-GPU product tests, then the weak / strong scaling lines
python -m pytest tests/test_multi_gpu.py tests/test_api_edges.py -m gpu -x -q 2>&1 | tail -2
scratch/scale.sh 2
