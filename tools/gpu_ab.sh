#!/bin/bash
# GPU visit: parity suites, then A/B of the range overlap (env switch) on the short bench, the default bench,
# the launch list and one ncu --set full capture.   usage: bash tools/gpu_ab.sh <tag> [notests]
TAG=${1:-s}; OUT=gpurun_out; mkdir -p $OUT
PYT="python -m pytest -m gpu -q -p no:cacheprovider --timeout 300 --timeout-method thread"
if [ "$2" != notests ]; then
  timeout -k 10 900 $PYT -x tests/test_gpu_parity.py > $OUT/t_parity_$TAG.log 2>&1; echo "parity rc=$?"; tail -n 3 $OUT/t_parity_$TAG.log
  timeout -k 10 1200 $PYT tests/test_gpu_configs.py > $OUT/t_configs_$TAG.log 2>&1; echo "configs-tests rc=$?"; tail -n 6 $OUT/t_configs_$TAG.log
fi
SHORT="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-configs"
for rep in 1 2; do
  WORDPIECE_B200_OVERLAP=0 timeout -k 10 300 $SHORT > $OUT/ab_${TAG}_serial_$rep.json 2> $OUT/ab_${TAG}_serial_$rep.err
  WORDPIECE_B200_OVERLAP=1 timeout -k 10 300 $SHORT > $OUT/ab_${TAG}_overlap_$rep.json 2> $OUT/ab_${TAG}_overlap_$rep.err
done
python - <<'PY'
import glob, json, os
for f in sorted(glob.glob("gpurun_out/ab_*_[12].json")):
    try:
        d = json.load(open(f))
    except Exception as e:
        print(f, "unreadable", e); continue
    r = d["roofline"]["kernel_ms_per_step"]
    print(os.path.basename(f), "%.3f ms  %.1f GB/s  K1 %.2f K2 %.2f K3 %.2f" % (d["ms_per_step"], d["value"], r["wp_split_kernel"], r["wp_match_kernel"], r["wp_scatter_kernel"]))
PY
timeout -k 10 900 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?"; tail -n 2 $OUT/bench_$TAG.err
SMALL="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-configs --mib 256"
timeout -k 10 300 $SMALL > $OUT/small_$TAG.json 2> $OUT/small_$TAG.err && \
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_$TAG.csv \
    $SMALL > $OUT/ncu_list_$TAG.log 2>&1; echo "ncu list rc=$?"
bash tools/ncu_full.sh $TAG en
