#!/bin/bash
# build variants of the library on the GPU box and time each: tools/gpu_variants.sh "<flags A>" "<flags B>" ...
for V in "$@"; do
  ORB_NVCC_EXTRA="$V" python visual-odometry-gpu_b200/build.py --force > /dev/null 2>&1 || { echo "build failed: $V"; continue; }
  python bench.py --steps 6 --warmup 3 --hot-only 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('[$V]', 'fps',round(d['value']),'stages',[round(x,3) for x in d['config']['stage_ms_per_step']],'e2e',round(d['e2e']['value']))"
done
python visual-odometry-gpu_b200/build.py --force > /dev/null 2>&1
