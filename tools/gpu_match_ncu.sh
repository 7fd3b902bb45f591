#!/bin/bash
# ncu capture of the tensor-core matcher (after the same command has exited 0 without ncu): tools/gpu_match_ncu.sh <tag>
TAG=${1:-m}; OUT=gpurun_out; mkdir -p $OUT
CMD="python tools/match_probe.py 149 2000"
$CMD > $OUT/match_plain_$TAG.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_match -c 4 -f -o $OUT/prof_match_$TAG $CMD > $OUT/ncu_match_$TAG.log 2>&1
echo "ncu rc=$?"; tail -2 $OUT/match_plain_$TAG.log
