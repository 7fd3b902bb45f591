#!/bin/bash
# round-2 first session: probes (instruction issue rates, TMA forms) + baseline tests + bench of HEAD
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/r2a_smi.txt 2>&1
timeout 120 tools/probes/pipe_probe > $OUT/r2a_pipe_probe.txt 2>&1; echo "pipe rc=$?"
timeout 120 tools/probes/tma_probe > $OUT/r2a_tma_probe.txt 2>&1; echo "tma rc=$?"
cat $OUT/r2a_pipe_probe.txt $OUT/r2a_tma_probe.txt
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > $OUT/r2a_tests.log; cat $OUT/r2a_tests.log
timeout 300 python bench.py --steps 10 --warmup 3 --hot-only > $OUT/r2a_bench.json 2> $OUT/r2a_bench.err; cat $OUT/r2a_bench.json | head -c 1500
