#!/bin/bash
# Is a parity failure flaky?  The batch suite N times with the product build, then the first failing test of the
# bounds build alone.   usage: bash tools/gpu_flaky.sh <tag> [n]
TAG=${1:-s}; N=${2:-8}; OUT=gpurun_out; mkdir -p $OUT
PYT="python -m pytest -m gpu -q -p no:cacheprovider --timeout 300 --timeout-method thread"
for i in $(seq 1 $N); do
  timeout -k 10 300 $PYT tests/test_gpu_batch.py > $OUT/t_flaky_${TAG}_$i.log 2>&1; echo "batch run $i rc=$? $(tail -n 1 $OUT/t_flaky_${TAG}_$i.log)"
done
timeout -k 10 600 $PYT tests/test_gpu_parity.py > $OUT/t_flaky_parity_$TAG.log 2>&1; echo "parity rc=$? $(tail -n 1 $OUT/t_flaky_parity_$TAG.log)"
export WORDPIECE_B200_LIB=$PWD/wordpiece_b200/lib/variants/libwordpiece_b200_bounds.so
timeout -k 10 300 $PYT -x tests/test_gpu_parity.py -k "multilingual_multi_tile" > $OUT/t_flaky_bounds_$TAG.log 2>&1; echo "bounds rc=$?"
grep -h "WP_CHECK failed" $OUT/t_flaky_bounds_$TAG.log | sed 's/block [0-9]*, thread [0-9]*//' | sort | uniq -c | head
unset WORDPIECE_B200_LIB
WORDPIECE_B200_TRACE=1 timeout -k 10 300 python tools/batch_trace.py > $OUT/batch_trace_$TAG.log 2>&1; grep -E "^rep|kernel" $OUT/batch_trace_$TAG.log
