#!/bin/bash
# Host-side sweeps: copy threads / streaming stores for the drop-in signature (interleaved, repeated), the batch entry.
TAG=${1:-s}; OUT=gpurun_out; mkdir -p $OUT
timeout -k 10 600 python tools/dropin_sweep.py > $OUT/dropin_sweep_$TAG.jsonl 2> $OUT/dropin_sweep_$TAG.err; echo "sweep rc=$?"; cat $OUT/dropin_sweep_$TAG.jsonl
for ss in 1 0 1 0; do
  echo "batch, stream stores $ss:"; WORDPIECE_B200_STREAM_STORES=$ss WORDPIECE_B200_TRACE=1 timeout -k 10 200 python tools/batch_trace.py 2>&1 | grep "^rep"
done
