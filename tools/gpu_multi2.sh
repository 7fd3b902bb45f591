#!/bin/bash
# 1080p and 4K (64 frames per GPU) lines at N GPUs: tools/gpu_multi2.sh <tag> <N>
TAG=$1; N=$2; OUT=gpurun_out; mkdir -p $OUT
for spec in "1080p --config 1080p" "4k_b64 --config 4k --frames 64"; do
  set -- $spec; name=$1; shift
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 "$@" \
      > $OUT/${TAG}_bench_${name}_${N}gpu.json 2> $OUT/${TAG}_bench_${name}_${N}gpu.err
  python -c "
import json
d=json.load(open('$OUT/${TAG}_bench_${name}_${N}gpu.json')); print('$name N=$N', 'fps', round(d['value']), 'e2e', round(d['e2e']['value']), 'link', round(d['e2e_link_bound']['value']))" || tail -5 $OUT/${TAG}_bench_${name}_${N}gpu.err
done
