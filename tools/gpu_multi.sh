#!/bin/bash
# multi-GPU bench lines (run under gpurun --gpus N): tools/gpu_multi.sh <tag> <N>
TAG=$1; N=$2; OUT=gpurun_out; mkdir -p $OUT
run() {  # name, extra args
  local name=$1; shift
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 "$@" \
      > $OUT/${TAG}_bench_${name}_${N}gpu.json 2> $OUT/${TAG}_bench_${name}_${N}gpu.err
  python - <<PY
import json
try:
    d=json.load(open("$OUT/${TAG}_bench_${name}_${N}gpu.json"))
    print("$name N=$N", "fps", round(d["value"]), "ms/step", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"]), "link", round(d["e2e_link_bound"]["value"]), "pass_frac", round(d["config"]["pass_hbm_frac"],4), "roof", round(d["roofline"]["frac"],4))
except Exception as e:
    print("$name failed", e); print(open("$OUT/${TAG}_bench_${name}_${N}gpu.err").read()[-800:])
PY
}
run kitti
if [ "$3" == "all" ]; then
  run 1080p --config 1080p
  for F in 16 32 64 128; do run 4k_b$F --config 4k --frames $F; done
fi
