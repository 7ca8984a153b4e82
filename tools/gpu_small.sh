#!/bin/bash
TAG=${1:-s}; OUT=gpurun_out; mkdir -p $OUT
for b in 4096 65536; do timeout -k 10 120 python tools/small_call.py --bytes $b 2>&1 | tail -n 1 | tee -a $OUT/small_call_$TAG.txt; done
timeout -k 10 200 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 200 --csv --log-file $OUT/small_call_launches_$TAG.csv python tools/small_call.py --calls 4 > $OUT/small_call_ncu_$TAG.log 2>&1; echo "ncu rc=$?"
python - <<PY
import csv
rows=list(csv.reader(open("$OUT/small_call_launches_$TAG.csv")))
h=None; out=[]
for r in rows:
    if 'Kernel Name' in r: h=r; continue
    if h and len(r)==len(h):
        d=dict(zip(h,r)); out.append((d['Kernel Name'][:34], d['Metric Value'], d['Grid Size']))
for o in out[-16:]: print(o)
PY
