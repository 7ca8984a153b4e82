#!/bin/bash
# GPU visit: all parity suites, the per-workload table (128 MiB each), the default bench.  usage: bash tools/gpu_check2.sh <tag> [nobench]
TAG=${1:-s}; OUT=gpurun_out; mkdir -p $OUT
PYT="python -m pytest -m gpu -q -p no:cacheprovider --timeout 300 --timeout-method thread"
timeout -k 10 900 $PYT -x tests/test_gpu_parity.py tests/test_gpu_batch.py > $OUT/t_parity_$TAG.log 2>&1; echo "parity+batch rc=$?"; tail -n 4 $OUT/t_parity_$TAG.log
timeout -k 10 1200 $PYT tests/test_gpu_configs.py > $OUT/t_configs_$TAG.log 2>&1; echo "configs-tests rc=$?"; tail -n 6 $OUT/t_configs_$TAG.log
for wl in en ru ja zh adv dirty open; do
  timeout -k 10 300 python tools/profile_workload.py $wl --mib 128 >> $OUT/wl_$TAG.jsonl 2>> $OUT/wl_$TAG.err
done
cat $OUT/wl_$TAG.jsonl
if [ "$2" != nobench ]; then
  timeout -k 10 900 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?"; tail -n 2 $OUT/bench_$TAG.err
fi
