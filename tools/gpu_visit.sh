#!/bin/bash
# One GPU-box visit: all three parity suites, the default bench (every leg), the start-up trace.
# usage: bash tools/gpu_visit.sh <tag> [ncu]
TAG=${1:-s}; OUT=gpurun_out; mkdir -p $OUT
PYT="python -m pytest -m gpu -q -p no:cacheprovider --timeout 300 --timeout-method thread"
timeout -k 10 900 $PYT -x tests/test_gpu_parity.py tests/test_gpu_batch.py > $OUT/t_parity_$TAG.log 2>&1; echo "parity+batch rc=$?"; tail -n 4 $OUT/t_parity_$TAG.log
timeout -k 10 1200 $PYT tests/test_gpu_configs.py > $OUT/t_configs_$TAG.log 2>&1; echo "configs-tests rc=$?"; tail -n 6 $OUT/t_configs_$TAG.log
timeout -k 10 900 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?"; tail -n 2 $OUT/bench_$TAG.err
timeout -k 10 300 python tools/startup_trace.py > $OUT/startup_$TAG.log 2>&1; echo "startup rc=$?"
if [ "$2" = ncu ]; then
  SMALL="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-configs --mib 256"
  timeout -k 10 300 $SMALL > $OUT/small_$TAG.json 2> $OUT/small_$TAG.err && \
  timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_$TAG.csv \
      $SMALL > $OUT/ncu_list_$TAG.log 2>&1; echo "ncu list rc=$?"
  bash tools/ncu_full.sh $TAG en
fi
