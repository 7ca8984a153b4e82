#!/bin/bash
# After a fix: the recorded fuzz seeds, the long-run repro, the suites in one process, the suites + fuzz under the bounds build.
TAG=${1:-s}; OUT=gpurun_out; mkdir -p $OUT
PYT="python -m pytest -m gpu -q -p no:cacheprovider --timeout 300 --timeout-method thread"
timeout -k 10 300 python tools/fuzz_gpu.py --seeds 1043,1140,1895 --repeat 3 > $OUT/verify_seeds_$TAG.log 2>&1; echo "seeds: $(tail -n 1 $OUT/verify_seeds_$TAG.log)"
timeout -k 10 300 python tools/long_repro.py 1140 > $OUT/verify_long_$TAG.log 2>&1; echo "long repro: $(tail -n 1 $OUT/verify_long_$TAG.log)"
for i in 1 2; do
  timeout -k 10 900 $PYT tests/test_gpu_parity.py tests/test_gpu_batch.py > $OUT/verify_suites_${TAG}_$i.log 2>&1; echo "parity+batch run $i: $(tail -n 1 $OUT/verify_suites_${TAG}_$i.log)"
done
timeout -k 10 1200 $PYT tests/test_gpu_configs.py > $OUT/verify_configs_$TAG.log 2>&1; echo "configs: $(tail -n 1 $OUT/verify_configs_$TAG.log)"
timeout -k 10 400 python tools/fuzz_gpu.py --seconds 240 > $OUT/verify_fuzz_$TAG.log 2>&1; echo "fuzz (product build): $(tail -n 1 $OUT/verify_fuzz_$TAG.log)"; grep -E "^FAIL" $OUT/verify_fuzz_$TAG.log | head -5
bash tools/gpu_bounds.sh $TAG
