#!/bin/bash
# One gpurun session: parity tests, smoke, bench (both arms), ncu launch list + full capture of the kernels.
# usage: tools/gpu_session.sh <tag>
TAG=${1:-r2}
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -8 > $OUT/tests_$TAG.log; tail -3 $OUT/tests_$TAG.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_ref_$TAG.json 2> $OUT/bench_$TAG.err; cat $OUT/bench_ref_$TAG.json | head -c 600; echo
timeout 900 python bench.py --steps 10 --warmup 3 > $OUT/bench_$TAG.json 2>> $OUT/bench_$TAG.err; head -c 1200 $OUT/bench_$TAG.json; echo; tail -3 $OUT/bench_$TAG.err
for c in 1080p 4k; do timeout 600 python bench.py --config $c --steps 10 --warmup 3 --no-ingest > $OUT/bench_${c}_$TAG.json 2> $OUT/bench_${c}_$TAG.err; head -c 300 $OUT/bench_${c}_$TAG.json; echo; done
SMALL="python bench.py --steps 2 --warmup 3 --frames 128 --hot-only"
$SMALL > $OUT/plain_$TAG.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_$TAG.csv $SMALL > $OUT/ncu_list_$TAG.log 2>&1
echo "ncu list rc=$?"
$SMALL > $OUT/plain2_$TAG.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_ -s 36 -c 6 -f -o $OUT/prof_$TAG $SMALL > $OUT/ncu_full_$TAG.log 2>&1
echo "ncu full rc=$?"
