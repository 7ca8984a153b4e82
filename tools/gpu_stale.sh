#!/bin/bash
# Stale-scratch hunt: the recorded fuzz seeds and the batch suite with the scratch re-poisoned before every range.
TAG=${1:-s}; OUT=gpurun_out; mkdir -p $OUT
PYT="python -m pytest -m gpu -q -p no:cacheprovider --timeout 300 --timeout-method thread"
export WORDPIECE_B200_POISON_SCRATCH=1
timeout -k 10 300 python tools/fuzz_gpu.py --seeds 1043,1140,1895 --repeat 1 > $OUT/stale_fuzz_$TAG.log 2>&1; cat $OUT/stale_fuzz_$TAG.log | cut -c1-400
timeout -k 10 300 $PYT tests/test_gpu_batch.py > $OUT/stale_batch_$TAG.log 2>&1; echo "batch: $(tail -n 1 $OUT/stale_batch_$TAG.log)"
timeout -k 10 600 $PYT tests/test_gpu_parity.py > $OUT/stale_parity_$TAG.log 2>&1; echo "parity: $(tail -n 1 $OUT/stale_parity_$TAG.log)"; grep -E "^FAILED" $OUT/stale_parity_$TAG.log | head
unset WORDPIECE_B200_POISON_SCRATCH
export WORDPIECE_B200_POISON=127
timeout -k 10 300 $PYT tests/test_gpu_batch.py > $OUT/stale_batch_p127_$TAG.log 2>&1; echo "batch, all allocations poisoned: $(tail -n 1 $OUT/stale_batch_p127_$TAG.log)"
