#!/bin/bash
# Which allocation's initial contents matter?  The batch suite and the recorded fuzz seeds under each poison class.
TAG=${1:-s}; OUT=gpurun_out; mkdir -p $OUT
PYT="python -m pytest -m gpu -q -p no:cacheprovider --timeout 300 --timeout-method thread"
for cls in 0 1 2 4 8 16 32 64; do
  export WORDPIECE_B200_POISON=$cls
  timeout -k 10 300 $PYT tests/test_gpu_batch.py > $OUT/poison_batch_${cls}_$TAG.log 2>&1; echo "class $cls batch: $(tail -n 1 $OUT/poison_batch_${cls}_$TAG.log)"
  timeout -k 10 300 python tools/fuzz_gpu.py --seeds 1043,1140,1895 --repeat 1 > $OUT/poison_fuzz_${cls}_$TAG.log 2>&1; echo "class $cls fuzz: $(tail -n 1 $OUT/poison_fuzz_${cls}_$TAG.log | cut -c1-60)"
done
