#!/bin/bash
# GPU visit: parity on the changed kernels, then launch lists + full captures of the dirty-web and ja workloads.
# usage: bash tools/gpu_profile2.sh <tag>
TAG=${1:-s}; OUT=gpurun_out; mkdir -p $OUT
PYT="python -m pytest -m gpu -q -p no:cacheprovider --timeout 300 --timeout-method thread"
timeout -k 10 900 $PYT -x tests/test_gpu_parity.py tests/test_gpu_batch.py > $OUT/t_parity_$TAG.log 2>&1; echo "parity+batch rc=$?"; tail -n 4 $OUT/t_parity_$TAG.log
timeout -k 10 1200 $PYT tests/test_gpu_configs.py -k "full_array or dirty or stress or sharded" > $OUT/t_configs_$TAG.log 2>&1; echo "configs-tests rc=$?"; tail -n 6 $OUT/t_configs_$TAG.log
for wl in en ru ja zh adv dirty open; do
  timeout -k 10 300 python tools/profile_workload.py $wl --mib 128 >> $OUT/wl_$TAG.jsonl 2>> $OUT/wl_$TAG.err
done
cat $OUT/wl_$TAG.jsonl
for wl in dirty ja; do
  timeout -k 10 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/launches_${wl}_$TAG.csv \
      python tools/profile_workload.py $wl --mib 128 --reps 1 > $OUT/ncu_list_${wl}_$TAG.log 2>&1; echo "ncu list $wl rc=$?"
  # one full range: seed + (2 MiB, 8 MiB) ranges = 1 + 2 x 4 launches to skip, then K1 K2 K2L K3 of the 64 MiB range
  timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:wp_ -s 9 -c 4 -f -o $OUT/prof_${wl}_$TAG \
      python tools/profile_workload.py $wl --mib 128 --reps 1 > $OUT/ncu_full_${wl}_$TAG.log 2>&1; echo "ncu full $wl rc=$?"
done
