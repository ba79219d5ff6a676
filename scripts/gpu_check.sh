#!/bin/bash
# Runs on the GPU box (under gpurun): GPU parity tests, frame-kernel phase timing, short bench.
# usage: scripts/gpu_check.sh <tag>
TAG=${1:-x}
OUT=gpurun_out
mkdir -p $OUT
python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/pytest_gpu_$TAG.log
python scripts/frame_timing.py > $OUT/frame_timing_$TAG.log 2>&1; echo "timing rc=$?"; tail -3 $OUT/frame_timing_$TAG.log
python bench.py --steps 5 --warmup 3 --batch 1024 --no-cpu --no-extras --e2e-steps 1 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?"
python - <<PY
import json
try:
    d=json.load(open("$OUT/bench_$TAG.json"))
    print("value", round(d["value"]), "fps  ms/step", round(d["ms_per_step"],3), "stage", d["stage_ms_per_step"], "pixel frac", round(d["roofline"]["frac"],3), "full frac", round(d["roofline"]["full_path_frac"],3), "e2e", round(d["e2e"]["value"]))
except Exception as e:
    print("bench parse failed", e)
PY
tail -3 $OUT/bench_$TAG.err
