#!/bin/bash
# GPU visit: the batch tests (new entry point), then A/B of lib/variants/*.so.  usage: bash tools/gpu_variants.sh <tag>
TAG=${1:-s}; OUT=gpurun_out; mkdir -p $OUT
PYT="python -m pytest -m gpu -q -p no:cacheprovider --timeout 300 --timeout-method thread"
timeout -k 10 600 $PYT -x tests/test_gpu_batch.py > $OUT/t_batch_$TAG.log 2>&1; echo "batch rc=$?"; tail -n 15 $OUT/t_batch_$TAG.log
bash tools/ab.sh $TAG
