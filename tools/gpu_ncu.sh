#!/bin/bash
# ncu --set full capture of kernels matching $2 on a short hot-only bench run; usage: tools/gpu_ncu.sh <tag> <regex> [skip] [count] [frames]
TAG=${1:-p}; RE=${2:-k_}; SKIP=${3:-3}; CNT=${4:-1}; FR=${5:-128}
OUT=gpurun_out; mkdir -p $OUT
SMALL="python bench.py --steps 2 --warmup 3 --frames $FR --hot-only"
$SMALL > $OUT/plain_$TAG.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:$RE -s $SKIP -c $CNT -f -o $OUT/prof_$TAG $SMALL > $OUT/ncu_full_$TAG.log 2>&1
echo "ncu rc=$?"; tail -3 $OUT/ncu_full_$TAG.log
