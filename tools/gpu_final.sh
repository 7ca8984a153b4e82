#!/bin/bash
# Final measurements of a round: default bench, reference arm, per-workload table, launch list, one ncu --set full
# capture (en, one full range), K2's L2 hit rate with warm caches on en and ja.   usage: bash tools/gpu_final.sh <tag>
TAG=${1:-s}; OUT=gpurun_out; mkdir -p $OUT
timeout -k 10 900 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?"; tail -n 2 $OUT/bench_$TAG.err
timeout -k 10 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_ref_$TAG.json 2> $OUT/bench_ref_$TAG.err; echo "reference arm rc=$?"
for wl in en ru ja zh adv dirty open; do
  timeout -k 10 300 python tools/profile_workload.py $wl --mib 128 >> $OUT/wl_$TAG.jsonl 2>> $OUT/wl_$TAG.err
done
cat $OUT/wl_$TAG.jsonl
SMALL="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-configs --mib 256"
timeout -k 10 300 $SMALL > $OUT/small_$TAG.json 2> $OUT/small_$TAG.err && \
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_$TAG.csv \
    $SMALL > $OUT/ncu_list_$TAG.log 2>&1; echo "ncu list rc=$?"
bash tools/ncu_full.sh $TAG en
for wl in en ja; do
  timeout -k 10 300 ncu --cache-control none --clock-control none -k regex:wp_match -s 2 -c 1 --csv \
      --metrics lts__t_sector_hit_rate.pct,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_read_lookup_hit.sum,lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum,l1tex__t_sector_hit_rate.pct,dram__bytes_read.sum,gpu__time_duration.sum \
      --log-file $OUT/k2_warm_l2_${wl}_$TAG.csv python tools/profile_workload.py $wl --mib 128 --reps 1 > $OUT/k2_warm_${wl}_$TAG.log 2>&1; echo "k2 warm-cache $wl rc=$?"
done
WORDPIECE_B200_TRACE=1 timeout -k 10 200 python tools/batch_trace.py > $OUT/batch_trace_$TAG.log 2>&1; echo "batch trace rc=$?"
WORDPIECE_B200_TRACE=2 timeout -k 10 200 python tools/batch_trace.py > $OUT/batch_stages_$TAG.log 2>&1; echo "batch stages rc=$?"
