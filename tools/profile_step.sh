#!/bin/bash
# ncu evidence for one training step of the headline configuration (on the GPU box, one GPU):
#   1. every launch of one eager step: device time, tensor-pipe activity, DRAM bytes  -> launches.csv
#   2. --set full of the two tensor-core samplers and of the hand-written GEMMs        -> *.ncu-rep
# bash tools/profile_step.sh <outdir under gpurun_out>
set -u
OUT=gpurun_out/${1:-prof}
mkdir -p $OUT
CMD="python bench.py --steps 1 --warmup 2 --no-graph --no-sweep --no-cpu-baseline"
# warm-up (2 steps) + graph-less bench legs launch ~900 kernels per step: skip the first 3 steps, take one step
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum \
    --clock-control none -s 2400 -c 800 --csv --log-file $OUT/launches.csv $CMD > $OUT/launches.log 2>&1
echo "launch list: exit $?"
timeout -s KILL 600 ncu --set full --import-source on --clock-control none -k regex:"sca_fwd_tc4|sca_bwd_tc2" -s 12 -c 2 -f -o $OUT/samplers $CMD > $OUT/samplers.log 2>&1
echo "samplers: exit $?"
timeout -s KILL 600 ncu --set full --clock-control none -k regex:gemm_tn_kernel -s 12 -c 2 -f -o $OUT/gemm $CMD > $OUT/gemm.log 2>&1
echo "gemm: exit $?"
ls -la $OUT
