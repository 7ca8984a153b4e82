#!/bin/bash
# Streaming-store staging copies: the micro-benchmark that motivates them, the batch entry with and without,
# parity of the batch / host-buffer entries, the extended fuzz sweep, the default bench.   usage: bash tools/gpu_stream.sh <tag>
TAG=${1:-s}; OUT=gpurun_out; mkdir -p $OUT
PYT="python -m pytest -m gpu -q -p no:cacheprovider --timeout 300 --timeout-method thread"
timeout -k 10 120 python tools/h2d_dirty_cache.py > $OUT/h2d_dirty_$TAG.txt 2>&1; cat $OUT/h2d_dirty_$TAG.txt
for ss in 0 1; do
  WORDPIECE_B200_STREAM_STORES=$ss WORDPIECE_B200_TRACE=1 timeout -k 10 200 python tools/batch_trace.py > $OUT/batch_trace_ss${ss}_$TAG.log 2>&1
  echo "stream stores $ss:"; grep "^rep" $OUT/batch_trace_ss${ss}_$TAG.log
done
WORDPIECE_B200_TRACE=2 timeout -k 10 200 python tools/batch_trace.py > $OUT/batch_stages_$TAG.log 2>&1; echo "batch stages rc=$?"
timeout -k 10 900 $PYT tests/test_gpu_parity.py tests/test_gpu_batch.py > $OUT/t_stream_$TAG.log 2>&1; echo "parity+batch: $(tail -n 1 $OUT/t_stream_$TAG.log)"
timeout -k 10 300 python tools/fuzz_gpu.py --seconds 120 --seed0 8000 > $OUT/fuzz_stream_$TAG.log 2>&1; echo "fuzz: $(tail -n 1 $OUT/fuzz_stream_$TAG.log)"; grep -E "^(FAIL|EXC)" $OUT/fuzz_stream_$TAG.log | head -5
timeout -k 10 900 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?"; tail -n 2 $OUT/bench_$TAG.err
python - <<PY
import json
d=json.load(open("$OUT/bench_$TAG.json"))
print("value", d["value"], "e2e", d["e2e"]["value"], "dropin", d["e2e_dropin"]["value"], d["e2e_dropin"]["seconds"])
for k in ("batch_10000x4KiB","batch_100000x256B"): print(k, d["configs"][k]["gb_per_s"], d["configs"][k]["seconds_best_of_5"])
PY
