#!/bin/bash
# compute-sanitizer memcheck + racecheck over DP and pipeline parity tests (hygiene)
mkdir -p gpurun_out
SEL_DP='small_problems or class_boundaries or tie_heavy or wildcards or trace_budget'
SEL_PIPE='c2_small_circular-c or c4_small_linear or fresh_31'
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 7 --print-limit 20 \
     python -m pytest tests/test_dp_parity.py tests/test_pipeline_parity.py tests/test_dp_kernel_variants.py -m gpu -x -q \
     -k "($SEL_DP) or ($SEL_PIPE) or teams_next or split_last" > gpurun_out/sanitizer_${tool}_r02.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|error" gpurun_out/sanitizer_${tool}_r02.log | tail -5
done
