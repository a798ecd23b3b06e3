#!/bin/bash
# One GPU-box visit: parity tests, bench, phase timers, ncu launch list + full capture of the sampler kernels.
# Usage (from the repo root, on the box): bash tools/gpu_round.sh <tag>
tag=${1:-run}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $out/pytest.log 2>&1; echo "pytest exit $?" >> $out/pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 > $out/bench.json 2> $out/bench.err; echo "bench exit $?" >> $out/bench.err
timeout 300 python tools/tc_timing.py --no-legacy > $out/tc_timing.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $out/ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sca_fwd_tc4_kernel -s 3 -c 1 -o $out/sca_fwd_tc4 \
    python tools/tc_timing.py --no-legacy > $out/ncu_full_fwd.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sca_bwd_tc2_kernel -s 2 -c 1 -o $out/sca_bwd_tc2 \
    python tools/tc_timing.py --no-legacy > $out/ncu_full_bwd.log 2>&1
tail -3 $out/pytest.log; cat $out/bench.json; tail -2 $out/bench.err; cat $out/tc_timing.txt
