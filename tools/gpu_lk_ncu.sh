#!/bin/bash
# batch tracker: timing for converging / non-converging tracks, then an ncu capture of k_lk_track (after the plain run exited 0)
TAG=${1:-lk}; OUT=gpurun_out; mkdir -p $OUT
python tools/lk_probe.py 129 0; python tools/lk_probe.py 129 1
CMD="python tools/lk_probe.py 17 1"
$CMD > $OUT/lk_plain_$TAG.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_lk_track -c 1 -f -o $OUT/prof_lk_$TAG $CMD > $OUT/ncu_lk_$TAG.log 2>&1
echo "ncu rc=$?"
