#!/bin/bash
# Round checkpoint on a B200 box: bench (both arms), the five configurations, and ONE ncu pass
# (launch list by default; "full" = one --set full capture of the three kernels of a whole 64 MiB range).
# usage: bash tools/checkpoint.sh <tag> [list|full]
TAG=${1:-cp}; MODE=${2:-list}; OUT=gpurun_out; mkdir -p $OUT
SMALL="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
if [ "$MODE" = list ]; then
  python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?"
  python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_ref_$TAG.json 2> $OUT/bench_ref_$TAG.err; echo "ref rc=$?"
  python tools/run_configs.py > $OUT/configs_$TAG.jsonl 2> $OUT/configs_$TAG.err; echo "configs rc=$?"
  $SMALL > $OUT/small_$TAG.json 2> $OUT/small_$TAG.err && \
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_$TAG.csv \
      $SMALL > $OUT/ncu_list_$TAG.log 2>&1; echo "ncu list rc=$?"
else
  $SMALL > $OUT/small_$TAG.json 2> $OUT/small_$TAG.err && \
  ncu --set full --clock-control none --import-source on -k regex:wp_ -s 6 -c 3 -f -o $OUT/prof_$TAG \
      $SMALL > $OUT/ncu_full_$TAG.log 2>&1; echo "ncu full rc=$?"
fi
