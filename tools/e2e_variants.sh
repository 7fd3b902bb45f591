#!/bin/bash
# e2e (host-staged) throughput for a list of staged wave sizes: tools/e2e_variants.sh 128 256 512
for C in "$@"; do
  python bench.py --steps 6 --warmup 3 --hot-only --chunk $C 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('chunk $C', 'value', round(d['value']), 'e2e', round(d['e2e']['value']))"
done
