#!/bin/bash
# quick GPU check: parity tests (optionally a -k filter) + hot-only bench; usage: tools/gpu_quick.sh <tag> [pytest -k expr]
TAG=${1:-q}; K=${2:-}
OUT=gpurun_out; mkdir -p $OUT
if [ -n "$K" ]; then timeout 900 python -m pytest tests -m gpu -x -q -k "$K" 2>&1 | tail -15 > $OUT/tests_$TAG.log; else timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > $OUT/tests_$TAG.log; fi
cat $OUT/tests_$TAG.log
timeout 300 python bench.py --steps 10 --warmup 3 --hot-only > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
python - <<PY
import json
try:
    d=json.load(open('$OUT/bench_$TAG.json'))
    print('fps',round(d['value']),'ms/step',round(d['ms_per_step'],3),'stages',[round(x,3) for x in d['config']['stage_ms_per_step']],'e2e',round(d['e2e']['value']),'clk',d['clocks']['sm_mhz'])
except Exception as e:
    print('bench failed',e); print(open('$OUT/bench_$TAG.err').read()[-1500:])
PY
