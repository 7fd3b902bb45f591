#!/bin/bash
# bench + ncu full capture (kernels matching $2, default all k_) ; usage: tools/gpu_prof.sh <tag> [regex] [skip] [count]
TAG=${1:-p}; RE=${2:-k_}; SKIP=${3:-48}; CNT=${4:-6}
OUT=gpurun_out; mkdir -p $OUT
python bench.py --steps 5 --warmup 3 --hot-only > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
python -c "
import json; d=json.load(open('$OUT/bench_$TAG.json')); print('fps',d['value'],'ms/step',d['ms_per_step'],'stages',d['config']['stage_ms_per_step'],'e2e',d['e2e']['value'],'roof',d['roofline']['frac'])"; tail -3 $OUT/bench_$TAG.err
SMALL="python bench.py --steps 2 --warmup 3 --frames 112 --hot-only"
$SMALL > $OUT/plain_$TAG.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:$RE -s $SKIP -c $CNT -f -o $OUT/prof_$TAG $SMALL > $OUT/ncu_full_$TAG.log 2>&1
echo "ncu rc=$?"
