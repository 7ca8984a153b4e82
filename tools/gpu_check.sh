#!/bin/bash
# One GPU-box visit: parity suite (bounded: a hung kernel must not hold the box), a short bench, the five
# configurations.  usage: bash tools/gpu_check.sh <tag> [quick|full]
TAG=${1:-x}; MODE=${2:-quick}; OUT=gpurun_out; mkdir -p $OUT
PYT="python -m pytest -m gpu -x -q -p no:cacheprovider --timeout 300 --timeout-method thread"
timeout -k 10 900 $PYT tests/test_gpu_parity.py > $OUT/t_parity_$TAG.log 2>&1; echo "parity rc=$?"; tail -3 $OUT/t_parity_$TAG.log
if [ "$MODE" = full ]; then
  timeout -k 10 1200 $PYT tests/test_gpu_configs.py > $OUT/t_configs_$TAG.log 2>&1; echo "configs-tests rc=$?"; tail -3 $OUT/t_configs_$TAG.log
fi
timeout -k 10 600 python bench.py --steps 5 --warmup 3 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?"; tail -2 $OUT/bench_$TAG.err
timeout -k 10 600 python tools/run_configs.py > $OUT/configs_$TAG.jsonl 2> $OUT/configs_$TAG.err; echo "configs rc=$?"; tail -2 $OUT/configs_$TAG.err
