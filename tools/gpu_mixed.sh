#!/bin/bash
# K3 mixed path (long segments kept out of the staging buffer): parity, fuzz (also under the bounds build), A/B on the
# dirty-web and English workloads against the build without it.   usage: bash tools/gpu_mixed.sh <tag>
TAG=${1:-s}; OUT=gpurun_out; mkdir -p $OUT
PYT="python -m pytest -m gpu -q -p no:cacheprovider --timeout 300 --timeout-method thread"
timeout -k 10 600 $PYT tests/test_gpu_parity.py tests/test_gpu_batch.py > $OUT/t_mixed_$TAG.log 2>&1; echo "parity+batch: $(tail -n 1 $OUT/t_mixed_$TAG.log)"
timeout -k 10 600 $PYT tests/test_gpu_configs.py -k "dirty or stress or sharded or full_array" > $OUT/t_mixed_cfg_$TAG.log 2>&1; echo "configs (dirty, stress, sharded, full arrays): $(tail -n 1 $OUT/t_mixed_cfg_$TAG.log)"
timeout -k 10 300 python tools/fuzz_gpu.py --seconds 100 --seed0 20000 > $OUT/fuzz_mixed_$TAG.log 2>&1; echo "fuzz: $(tail -n 1 $OUT/fuzz_mixed_$TAG.log)"; grep -E "^(FAIL|EXC)" $OUT/fuzz_mixed_$TAG.log | head -5
WORDPIECE_B200_LIB=$PWD/wordpiece_b200/lib/variants/libwordpiece_b200_bounds.so timeout -k 10 300 python tools/fuzz_gpu.py --seconds 60 --seed0 23000 > $OUT/fuzz_mixed_bounds_$TAG.log 2>&1; echo "bounds fuzz: $(tail -n 1 $OUT/fuzz_mixed_bounds_$TAG.log)"
grep -h "WP_CHECK failed" $OUT/fuzz_mixed_bounds_$TAG.log | sort | uniq -c | head -5
for lib in default nomixed default nomixed; do
  if [ "$lib" = default ]; then unset WORDPIECE_B200_LIB; else export WORDPIECE_B200_LIB=$PWD/wordpiece_b200/lib/variants/libwordpiece_b200_$lib.so; fi
  for wl in dirty en adv; do echo -n "$lib " >> $OUT/ab_mixed_$TAG.txt; timeout -k 10 200 python tools/profile_workload.py $wl --mib 128 >> $OUT/ab_mixed_$TAG.txt 2>> $OUT/ab_mixed_$TAG.err; done
done
unset WORDPIECE_B200_LIB
cut -c1-260 $OUT/ab_mixed_$TAG.txt
