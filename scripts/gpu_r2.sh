#!/bin/bash
# Runs on the GPU box (under gpurun): GPU parity tests, then the default bench and the reference arm.
# usage: scripts/gpu_r2.sh <tag> [pytest-args]
TAG=${1:-x}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi -L > $OUT/gpus_$TAG.txt 2>&1; nproc >> $OUT/gpus_$TAG.txt
timeout 1500 python -m pytest tests -m gpu -q -x ${2:-} > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -15 $OUT/pytest_gpu_$TAG.log
timeout 600 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?"; tail -3 $OUT/bench_$TAG.err
timeout 600 python bench.py --impl reference > $OUT/bench_ref_$TAG.json 2> $OUT/bench_ref_$TAG.err; echo "bench ref rc=$?"; tail -3 $OUT/bench_ref_$TAG.err
python - <<PY
import json
try:
    d=json.load(open("$OUT/bench_$TAG.json"))
    print("value", round(d["value"]), "fps  ms/step", round(d["ms_per_step"],3), "stage", {k: round(v,3) for k,v in d["stage_ms_per_step"].items()}, "pixel frac", round(d["roofline"]["frac"],3), "full frac", round(d["roofline"]["full_path_frac"],3), "e2e", round(d["e2e"]["value"]), "nomask", round(d["e2e"]["without_mask_download"]))
    print("cpu", d["cpu_baseline"])
    print("extras", json.dumps(d["extras"])[:3000])
    r=json.load(open("$OUT/bench_ref_$TAG.json"))
    print("ref", r["value"], r["cpu_baseline"])
except Exception as e:
    print("bench parse failed", e)
PY
