#!/bin/bash
# GPU visit: all parity suites (product build), the same under the bounds-checked build + fuzz, the default bench.
# usage: bash tools/gpu_full.sh <tag>
TAG=${1:-s}; OUT=gpurun_out; mkdir -p $OUT
PYT="python -m pytest -m gpu -q -p no:cacheprovider --timeout 300 --timeout-method thread"
timeout -k 10 900 $PYT -x tests/test_gpu_parity.py tests/test_gpu_batch.py > $OUT/t_parity_$TAG.log 2>&1; echo "parity+batch rc=$?"; tail -n 3 $OUT/t_parity_$TAG.log
timeout -k 10 1200 $PYT tests/test_gpu_configs.py > $OUT/t_configs_$TAG.log 2>&1; echo "configs-tests rc=$?"; tail -n 4 $OUT/t_configs_$TAG.log
bash tools/gpu_bounds.sh $TAG
timeout -k 10 900 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?"; tail -n 2 $OUT/bench_$TAG.err
