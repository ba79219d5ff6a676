#!/bin/bash
# Round-2 evidence run on the GPU box (under gpurun): plain bench first, then the ncu launch list of the same command and one
# `--set full` capture per kernel (each only after the plain command has exited 0), then the Bayer and stress captures.
set -u
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
BENCH_SMALL="python bench.py --steps 2 --warmup 3 --batch 1024 --no-cpu --no-extras --e2e-steps 1"
$BENCH_SMALL > $OUT/bench_small_$TAG.json 2> $OUT/bench_small_$TAG.err || { echo "plain small bench failed"; tail -5 $OUT/bench_small_$TAG.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches_$TAG.csv $BENCH_SMALL > $OUT/ncu_launches_$TAG.log 2>&1
echo "ncu launches rc=$?"
for k in pixel_bgr_kernel emit_kernel label_kernel contour_kernel fit_kernel order_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -f -o $OUT/${k}_$TAG $BENCH_SMALL > $OUT/ncu_full_${k}_$TAG.log 2>&1
  echo "ncu full $k rc=$?"
done
python scripts/bayer_bench.py > $OUT/bayer_bench_$TAG.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:bayer_strip_kernel -s 8 -c 1 -f -o $OUT/bayer_strip_kernel_$TAG python scripts/bayer_bench.py > $OUT/ncu_full_bayer_$TAG.log 2>&1
echo "ncu bayer rc=$?"
python scripts/stress_bench.py 16 > $OUT/stress_bench_$TAG.log 2>&1 && \
for k in label_kernel order_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 4 -c 1 -f -o $OUT/stress_${k}_$TAG python scripts/stress_bench.py 16 > $OUT/ncu_stress_${k}_$TAG.log 2>&1
  echo "ncu stress $k rc=$?"
done
python scripts/shim_timing.py 2>&1 | tail -1
python scripts/latency_bench.py 2>&1 | tail -1
ls -la $OUT | tail -25
