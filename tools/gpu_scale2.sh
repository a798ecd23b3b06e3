#!/bin/bash
# 2-GPU visit: N=1 and N=2 bench lines on the same box, N=2 with the gradient exchange as one exposed all-reduce
# (VER_BUCKET_BYTES=0) and as overlapped buckets (default); the 2-rank NCCL parity test.
tag=${1:-scale2}
N=${2:-2}
out=gpurun_out/$tag
mkdir -p $out
B="--steps 20 --warmup 3 --no-sweep --no-cpu-baseline"
timeout 300 python bench.py --gpus 1 $B > $out/bench_n1.json 2> $out/bench_n1.err
for bb in 0 8388608; do
  VER_BUCKET_BYTES=$bb timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
      bench.py --gpus $N $B > $out/bench_n${N}_bucket$bb.json 2> $out/bench_n${N}_bucket$bb.err
done
timeout 400 python -m pytest tests/test_gpu_ddp.py -m gpu -q -x > $out/pytest_ddp.log 2>&1
python - <<PY
import json
for f in ("bench_n1", "bench_n${N}_bucket0", "bench_n${N}_bucket8388608"):
    try:
        d = json.load(open("$out/" + f + ".json"))
        print(f, d["value"], d["ms_per_step"], d["e2e"]["value"], d["config"].get("execution"))
    except Exception as e:
        print(f, "FAILED", e)
PY
tail -2 $out/pytest_ddp.log
