#!/bin/bash
# One `ncu --set full` capture of the kernels of one FULL 64 MiB range (K1, K2, K2L, K3) of the English workload.
# Launch order of a call: word-table seed, then K1 K2 K2L K3 per range (2 MiB, 8 MiB, 64 MiB, ...): skip 9, take 4.
# usage: bash tools/ncu_full.sh <tag> [workload]
TAG=${1:-x}; WL=${2:-en}; OUT=gpurun_out; mkdir -p $OUT
SMALL="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-configs --workload $WL --mib 256"
timeout -k 10 300 $SMALL > $OUT/small_$TAG.json 2> $OUT/small_$TAG.err && \
timeout -k 10 900 ncu --set full --clock-control none --import-source on -k regex:wp_ -s 9 -c 4 -f -o $OUT/prof_$TAG \
    $SMALL > $OUT/ncu_full_$TAG.log 2>&1; echo "ncu full rc=$?"
