#!/bin/bash
# matcher check on the GPU box: parity tests of the matcher, then the probe for the shipped build and for each variant given
# usage: tools/gpu_match.sh "<nvcc flags of variant A>" ...
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_match.py -m gpu -x -q 2>&1 | tail -5
timeout 300 python tools/match_probe.py
for V in "$@"; do
  D=$(mktemp -d)
  ORB_NVCC_EXTRA="$V" python visual-odometry-gpu_b200/build.py --force --out $D/liborb_b200.so > $D/build.log 2>&1 || { echo "build failed: $V"; tail -5 $D/build.log; continue; }
  echo "[$V]"; ORB_B200_LIB=$D/liborb_b200.so timeout 300 python tools/match_probe.py
done
