#!/bin/bash
# Round-2 visit: full parity suite, smoke, bench (default line), then the ncu evidence of profile_step.sh.
tag=${1:-run}
out=gpurun_out/$tag
mkdir -p $out
timeout 420 python -m pytest tests -m gpu -q -x > $out/pytest.log 2>&1; echo "pytest exit $?" >> $out/pytest.log
timeout 120 python __graft_entry__.py smoke > $out/smoke.log 2>&1; echo "smoke exit $?" >> $out/smoke.log
timeout 400 python bench.py --steps 20 --warmup 3 > $out/bench.json 2> $out/bench.err; echo "bench exit $?" >> $out/bench.err
tail -4 $out/pytest.log; tail -2 $out/smoke.log; cat $out/bench.json; tail -1 $out/bench.err
if [ "${2:-}" = "prof" ]; then bash tools/profile_step.sh $tag; fi
if [ "${2:-}" = "bwdprof" ]; then
  timeout -s KILL 600 ncu --set full --import-source on --clock-control none -k regex:"sca_bwd_tc2" -s 3 -c 1 -f -o $out/bwd \
      python bench.py --steps 1 --warmup 2 --no-graph --no-sweep --no-cpu-baseline > $out/bwd_ncu.log 2>&1
  echo "bwd ncu: exit $?"
fi
if [ "${2:-}" = "fwdprof" ]; then
  timeout -s KILL 600 ncu --set full --import-source on --clock-control none -k regex:"sca_fwd_tc4" -s 3 -c 1 -f -o $out/fwd \
      python bench.py --steps 1 --warmup 2 --no-graph --no-sweep --no-cpu-baseline > $out/fwd_ncu.log 2>&1
  echo "fwd ncu: exit $?"
fi
