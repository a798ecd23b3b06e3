#!/bin/bash
# compute-sanitizer over the tensor-core sampler kernels (SURVEY section 5 aux row).
# Smallest parity case of every kernel generation: forward tc3 / tc4 / tc5 / tc6 / tc7 / block, backward tc2 / tc.
#   bash tools/sanitize.sh   (on the GPU box; writes gpurun_out/sanitizer/*.txt)
set -u
mkdir -p gpurun_out/sanitizer
SEL='tc_sampler_forward_backward_vs_oracle and 32-grid1-2'
for tool in memcheck racecheck synccheck; do
  timeout -s KILL 600 compute-sanitizer --tool $tool --print-limit 20 \
    python -m pytest tests/test_gpu_parity.py -q -x -k "$SEL" > gpurun_out/sanitizer/$tool.txt 2>&1
  echo "== $tool: exit $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|hazard" gpurun_out/sanitizer/$tool.txt | tail -5
done
