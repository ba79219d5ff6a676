#!/bin/bash
# Runs on the GPU box (under gpurun): Bayer pixel-stage timing, then one ncu --set full capture of the Bayer kernel.
# usage: scripts/gpu_bayer_profile.sh <tag>
TAG=${1:-x}
OUT=gpurun_out
mkdir -p $OUT
python scripts/bayer_bench.py > $OUT/bayer_bench_$TAG.log 2>&1 || { echo "bayer bench failed"; tail -5 $OUT/bayer_bench_$TAG.log; exit 1; }
tail -2 $OUT/bayer_bench_$TAG.log
ncu --set full --clock-control none --import-source on -k regex:bayer -s 4 -c 1 -f -o $OUT/pixel_bayer_$TAG python scripts/bayer_bench.py > $OUT/ncu_bayer_$TAG.log 2>&1
echo "ncu bayer rc=$?"
