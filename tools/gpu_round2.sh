#!/bin/bash
# One GPU-box visit for the state after the decoder work: full parity suite, bench (train + the batch-64
# inference configuration), ncu launch list of the bench, ncu --set full of the 3-D sampler kernels.
# Usage (from the repo root, on the box): bash tools/gpu_round2.sh <tag>
tag=${1:-run}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/smi.txt 2>&1
timeout 600 python -m pytest tests -m gpu -q > $out/pytest.log 2>&1; echo "pytest exit $?" >> $out/pytest.log
timeout 300 python bench.py --steps 10 --warmup 3 > $out/bench.json 2> $out/bench.err; echo "bench exit $?" >> $out/bench.err
timeout 240 python bench.py --mode infer --batch 64 --steps 10 --warmup 3 --no-cpu-baseline > $out/bench_infer_b64.json 2> $out/bench_infer.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $out/ncu_bench.log 2>&1
timeout 150 ncu --set full --clock-control none --import-source on -k regex:msda3d_fwd_kernel -s 13 -c 1 -o $out/msda3d_fwd \
    python tools/msda3d_bench.py > $out/ncu_msda3d_fwd.log 2>&1
timeout 150 ncu --set full --clock-control none --import-source on -k regex:msda3d_bwd_kernel -s 13 -c 1 -o $out/msda3d_bwd \
    python tools/msda3d_bench.py > $out/ncu_msda3d_bwd.log 2>&1
tail -4 $out/pytest.log; cat $out/bench.json; tail -2 $out/bench.err; cat $out/bench_infer_b64.json; tail -2 $out/bench_infer.err
