#!/bin/bash
# What the driver runs at round end (all GPU tests, smoke, default bench) + the fuzz sweep under both builds.
TAG=${1:-s}; OUT=gpurun_out; mkdir -p $OUT
timeout -k 10 1200 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider > $OUT/t_all_$TAG.log 2>&1; echo "pytest -m gpu: $(tail -n 1 $OUT/t_all_$TAG.log)"
timeout -k 10 200 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke_$TAG.log 2>&1; echo "smoke: $(tail -n 1 $OUT/smoke_$TAG.log)"
timeout -k 10 300 python tools/fuzz_gpu.py --seconds 120 --seed0 12000 > $OUT/fuzz_$TAG.log 2>&1; echo "fuzz: $(tail -n 1 $OUT/fuzz_$TAG.log)"; grep -E "^(FAIL|EXC)" $OUT/fuzz_$TAG.log | head -5
WORDPIECE_B200_LIB=$PWD/wordpiece_b200/lib/variants/libwordpiece_b200_bounds.so timeout -k 10 300 python tools/fuzz_gpu.py --seconds 90 --seed0 15000 > $OUT/fuzz_bounds_$TAG.log 2>&1; echo "bounds fuzz: $(tail -n 1 $OUT/fuzz_bounds_$TAG.log)"
grep -h "WP_CHECK failed" $OUT/fuzz_bounds_$TAG.log | sort | uniq -c | head -5
timeout -k 10 900 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?"; tail -n 2 $OUT/bench_$TAG.err
python - <<PY
import json
d=json.load(open("$OUT/bench_$TAG.json"))
print("value", d["value"], "frac", d["roofline"]["frac"], "traffic", d["roofline"]["traffic"], "e2e", d["e2e"]["value"], "dropin", d["e2e_dropin"]["value"])
for k in ("batch_10000x4KiB","batch_100000x256B"): print(k, d["configs"][k]["gb_per_s"])
PY
