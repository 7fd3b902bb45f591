#!/bin/bash
# ncu launch list (per-launch device time) of a short hot-only bench run; usage: tools/gpu_list.sh <tag> [frames]
TAG=${1:-l}; FR=${2:-128}
OUT=gpurun_out; mkdir -p $OUT
SMALL="python bench.py --steps 2 --warmup 3 --frames $FR --hot-only"
$SMALL > $OUT/plain_$TAG.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_$TAG.csv $SMALL > $OUT/ncu_list_$TAG.log 2>&1
echo "ncu list rc=$?"
python - <<PY
import csv, collections
rows=[r for r in csv.reader(open('$OUT/launches_$TAG.csv')) if len(r)>5]
h=rows[0]; ki=h.index('Kernel Name'); vi=h.index('Metric Value')
d=collections.defaultdict(list)
for r in rows[1:]:
    try: d[r[ki].split('(')[0]].append(float(r[vi].replace(',','')))
    except: pass
for k,v in d.items(): print('%-40s n=%3d avg %.1f us min %.1f'%(k,len(v),sum(v)/len(v)/1000 if max(v)>1e4 else sum(v)/len(v),min(v)))
PY
