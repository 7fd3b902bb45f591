#!/bin/bash
# One gpurun session: parity tests, smoke, bench (both arms), ncu launch list + full capture of the kernels.
# usage: tools/gpu_session.sh <tag>
TAG=${1:-r1}
OUT=gpurun_out
mkdir -p $OUT
python -m pytest tests -m gpu -q 2>&1 | tail -8 > $OUT/tests_$TAG.log; tail -3 $OUT/tests_$TAG.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_ref_$TAG.json 2> $OUT/bench_$TAG.err; cat $OUT/bench_ref_$TAG.json
python bench.py --steps 10 --warmup 3 > $OUT/bench_$TAG.json 2>> $OUT/bench_$TAG.err; cat $OUT/bench_$TAG.json; tail -5 $OUT/bench_$TAG.err
SMALL="python bench.py --steps 2 --warmup 3 --frames 128 --hot-only"
$SMALL > $OUT/plain_$TAG.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_$TAG.csv $SMALL > $OUT/ncu_list_$TAG.log 2>&1
echo "ncu list rc=$?"
$SMALL > $OUT/plain2_$TAG.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_ -s 30 -c 5 -f -o $OUT/prof_$TAG $SMALL > $OUT/ncu_full_$TAG.log 2>&1
echo "ncu full rc=$?"
