#!/bin/bash
# compute-sanitizer on single parity cases (by pytest node id suffix): bash tools/sanitize_one.sh <tool> <case> ...
tool=$1; shift
mkdir -p gpurun_out/sanitizer2
for c in "$@"; do
  timeout -s KILL 400 compute-sanitizer --tool $tool --print-limit 4 \
    python -m pytest "tests/test_gpu_parity.py::test_tc_sampler_forward_backward_vs_oracle[$c]" -q -x > gpurun_out/sanitizer2/${tool}_$c.txt 2>&1
  echo "== $tool $c: exit $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|Missing|Race reported" gpurun_out/sanitizer2/${tool}_$c.txt | cut -c1-150 | head -6
done
