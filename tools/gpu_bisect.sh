#!/bin/bash
# Which build first fails the recorded fuzz seeds?  usage: bash tools/gpu_bisect.sh <tag> <seeds> [repeat]
TAG=${1:-s}; SEEDS=${2:-1043,1140,1895}; REP=${3:-4}; OUT=gpurun_out; mkdir -p $OUT
for lib in wordpiece_b200/lib/variants/libwordpiece_b200_${BISECT_GLOB:-c*}.so default; do
  unset WORDPIECE_B200_LIB WORDPIECE_B200_POISON
  case $lib in
    default) name=HEAD;;
    poison) name=HEAD+poison; export WORDPIECE_B200_POISON=1;;
    bounds) name=bounds; export WORDPIECE_B200_LIB=$PWD/wordpiece_b200/lib/variants/libwordpiece_b200_bounds.so;;
    *) name=$(basename $lib .so); name=${name#libwordpiece_b200_}; export WORDPIECE_B200_LIB=$PWD/$lib;;
  esac
  timeout -k 10 600 python tools/fuzz_gpu.py --seeds $SEEDS --repeat $REP > $OUT/bisect_${TAG}_$name.log 2>&1
  echo "== $name: $(tail -n 1 $OUT/bisect_${TAG}_$name.log)"; grep -E "^FAIL|sizes|WP_CHECK" $OUT/bisect_${TAG}_$name.log | sort | uniq -c | sort -rn | head -8
done
