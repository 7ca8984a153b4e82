#!/bin/bash
# A/B of tuning variants on the GPU box: the default library and every lib/variants/*.so through the same
# short bench (device-resident 1 GiB English, kernel times).  usage: bash tools/ab.sh <tag> [extra bench args]
TAG=${1:-ab}; shift; OUT=gpurun_out; mkdir -p $OUT
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-configs $@"
for rep in 1 2; do
  timeout -k 10 300 $B > $OUT/ab_${TAG}_default_$rep.json 2> $OUT/ab_${TAG}_default_$rep.err
  for lib in wordpiece_b200/lib/variants/libwordpiece_b200_*.so; do
    [ -e "$lib" ] || continue
    name=$(basename $lib .so); name=${name#libwordpiece_b200_}
    [ "$name" = bounds ] && continue   # the debug build is for tools/gpu_bounds.sh
    WORDPIECE_B200_LIB=$PWD/$lib timeout -k 10 300 $B > $OUT/ab_${TAG}_${name}_$rep.json 2> $OUT/ab_${TAG}_${name}_$rep.err
  done
done
python - <<'PY'
import glob, json, os
rows = {}
for f in sorted(glob.glob("gpurun_out/ab_*_[12].json")):
    try:
        d = json.load(open(f))
    except Exception:
        continue
    k = os.path.basename(f)[:-7]
    r = d["roofline"]["kernel_ms_per_step"]
    rows.setdefault(k, []).append((d["ms_per_step"], r["wp_split_kernel"], r["wp_match_kernel"], r["wp_scatter_kernel"]))
for k, v in rows.items():
    print(k, " | ".join("%.3f (%.2f %.2f %.2f)" % x for x in v))
PY
