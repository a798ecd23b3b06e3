#!/bin/bash
# Final visit of the round: full parity suite, smoke, bench, up_sample and shipped-head measurements.
tag=${1:-run}
out=gpurun_out/$tag
mkdir -p $out
timeout 300 python -m pytest tests -m gpu -q > $out/pytest.log 2>&1; echo "pytest exit $?" >> $out/pytest.log
timeout 120 python __graft_entry__.py smoke > $out/smoke.log 2>&1; echo "smoke exit $?" >> $out/smoke.log
timeout 120 python tools/upsample_bench.py > $out/upsample_bench.txt 2>&1
timeout 180 python tools/vocc_shipped_bench.py > $out/vocc_shipped_bench.txt 2>&1
timeout 300 python bench.py --steps 10 --warmup 3 > $out/bench.json 2> $out/bench.err; echo "bench exit $?" >> $out/bench.err
tail -4 $out/pytest.log; tail -2 $out/smoke.log; cat $out/upsample_bench.txt; grep -v Warn $out/vocc_shipped_bench.txt | tail -6; cat $out/bench.json; tail -1 $out/bench.err
