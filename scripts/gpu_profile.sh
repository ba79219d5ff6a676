#!/bin/bash
# Runs on the GPU box (under gpurun): plain bench first, then the ncu launch list and one full capture of a kernel.
# usage: scripts/gpu_profile.sh <tag> <kernel-regex> [skip]
set -u
TAG=${1:-r01}
KREGEX=${2:-pixel_bgr}
SKIP=${3:-8}
OUT=gpurun_out
mkdir -p $OUT
BENCH_SMALL="python bench.py --steps 2 --warmup 3 --batch 256 --no-cpu --e2e-steps 1"
$BENCH_SMALL > $OUT/bench_small_$TAG.json 2> $OUT/bench_small_$TAG.err || { echo "plain small bench failed"; tail -5 $OUT/bench_small_$TAG.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches_$TAG.csv $BENCH_SMALL > $OUT/ncu_launches_$TAG.log 2>&1
echo "ncu launches rc=$?"
ncu --set full --clock-control none --import-source on -k regex:$KREGEX -s $SKIP -c 2 -o $OUT/${KREGEX}_$TAG $BENCH_SMALL > $OUT/ncu_full_$TAG.log 2>&1
echo "ncu full rc=$?"
ls -la $OUT | tail -8
