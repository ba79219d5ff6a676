#!/bin/bash
# Runs on the GPU box (under gpurun): plain bench first, then the ncu launch list and one full capture per kernel.
# usage: scripts/gpu_profile.sh <tag>
set -u
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
BENCH_SMALL="python bench.py --steps 2 --warmup 3 --batch 1024 --no-cpu --no-extras --e2e-steps 1"
$BENCH_SMALL > $OUT/bench_small_$TAG.json 2> $OUT/bench_small_$TAG.err || { echo "plain small bench failed"; tail -5 $OUT/bench_small_$TAG.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches_$TAG.csv $BENCH_SMALL > $OUT/ncu_launches_$TAG.log 2>&1
echo "ncu launches rc=$?"
for k in pixel_bgr_kernel emit_kernel label_kernel contour_kernel fit_kernel order_kernel; do  # regex: templates match by prefix
  ncu --set full --clock-control none --import-source on -k regex:$k -s 6 -c 1 -f -o $OUT/${k}_$TAG $BENCH_SMALL > $OUT/ncu_full_${k}_$TAG.log 2>&1
  echo "ncu full $k rc=$?"
done
ls -la $OUT | tail -12
