#!/bin/bash
# e2e throughput for a list of staged wave sizes (device wave size unchanged): tools/staged_sweep.sh <preset> 32 48 64 ...  (0 = the default)
P=$1; shift
for C in "$@"; do
  ORB_B200_STAGED_WAVE=$C python bench.py --config $P --steps 8 --warmup 3 --hot-only 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$P staged $C', 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'link', round(d['e2e_link_bound']['value']))"
done
