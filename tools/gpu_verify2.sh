#!/bin/bash
TAG=${1:-s}; OUT=gpurun_out; mkdir -p $OUT
PYT="python -m pytest -m gpu -q -p no:cacheprovider --timeout 300 --timeout-method thread"
timeout -k 10 300 python tools/fuzz_gpu.py --seeds 1043,1140,1895,2149,2559,2773 --repeat 2 > $OUT/verify_seeds_$TAG.log 2>&1; echo "seeds: $(tail -n 1 $OUT/verify_seeds_$TAG.log)"
timeout -k 10 300 python tools/long_repro.py 1140 > $OUT/verify_long_$TAG.log 2>&1; echo "long repro: $(tail -n 1 $OUT/verify_long_$TAG.log)"
timeout -k 10 900 $PYT tests/test_gpu_parity.py tests/test_gpu_batch.py > $OUT/verify_suites_${TAG}.log 2>&1; echo "parity+batch: $(tail -n 1 $OUT/verify_suites_${TAG}.log)"
timeout -k 10 300 python tools/fuzz_gpu.py --seconds 150 > $OUT/verify_fuzz_$TAG.log 2>&1; echo "fuzz (product build): $(tail -n 1 $OUT/verify_fuzz_$TAG.log)"; grep -E "^FAIL" $OUT/verify_fuzz_$TAG.log | head -5
export WORDPIECE_B200_LIB=$PWD/wordpiece_b200/lib/variants/libwordpiece_b200_bounds.so
timeout -k 10 900 $PYT tests/test_gpu_parity.py tests/test_gpu_batch.py > $OUT/verify_bounds_${TAG}.log 2>&1; echo "bounds parity+batch: $(tail -n 1 $OUT/verify_bounds_${TAG}.log)"
timeout -k 10 300 python tools/fuzz_gpu.py --seconds 120 --seed0 5000 > $OUT/verify_bounds_fuzz_$TAG.log 2>&1; echo "bounds fuzz: $(tail -n 1 $OUT/verify_bounds_fuzz_$TAG.log)"
grep -h "WP_CHECK failed" $OUT/verify_bounds_${TAG}.log $OUT/verify_bounds_fuzz_$TAG.log | sed 's/block [0-9]*, thread [0-9]*//' | sort | uniq -c | head
