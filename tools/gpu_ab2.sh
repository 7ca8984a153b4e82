#!/bin/bash
# GPU visit: batch trace, A/B of lib/variants on the ja / en / dirty workloads (128 MiB) and on the 1 GiB bench,
# then the parity suites and the fuzz sweep under the bounds-checked build.   usage: bash tools/gpu_ab2.sh <tag>
TAG=${1:-s}; OUT=gpurun_out; mkdir -p $OUT
PYT="python -m pytest -m gpu -q -p no:cacheprovider --timeout 300 --timeout-method thread"
timeout -k 10 900 $PYT -x tests/test_gpu_parity.py tests/test_gpu_batch.py > $OUT/t_parity_$TAG.log 2>&1; echo "parity+batch rc=$?"; tail -n 3 $OUT/t_parity_$TAG.log
WORDPIECE_B200_TRACE=1 timeout -k 10 300 python tools/batch_trace.py > $OUT/batch_trace_$TAG.log 2>&1; echo "batch trace rc=$?"; grep "^rep" $OUT/batch_trace_$TAG.log
for lib in default wordpiece_b200/lib/variants/libwordpiece_b200_*.so; do
  if [ "$lib" = default ]; then name=default; unset WORDPIECE_B200_LIB; else
    name=$(basename $lib .so); name=${name#libwordpiece_b200_}; [ "$name" = bounds ] && continue; export WORDPIECE_B200_LIB=$PWD/$lib; fi
  for wl in ja en dirty adv; do
    echo -n "$name " >> $OUT/abwl_$TAG.txt
    timeout -k 10 300 python tools/profile_workload.py $wl --mib 128 >> $OUT/abwl_$TAG.txt 2>> $OUT/abwl_$TAG.err
  done
done
unset WORDPIECE_B200_LIB
cat $OUT/abwl_$TAG.txt
bash tools/ab.sh $TAG
bash tools/gpu_bounds.sh $TAG
