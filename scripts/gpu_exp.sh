#!/bin/bash
# experiment runner: each "run TAG ENV..." runs the short bench with that environment
OUT=gpurun_out; mkdir -p $OUT
run() { tag=$1; shift; env "$@" python bench.py --steps 5 --warmup 3 --batch 1024 --no-cpu --e2e-steps 1 > $OUT/exp_$tag.json 2> $OUT/exp_$tag.err
  python - <<PY
import json
try:
    d=json.load(open("$OUT/exp_$tag.json"))
    print("$tag: value", round(d["value"]), "ms/step", round(d["ms_per_step"],3), "stage", {k:round(v,3) for k,v in d["stage_ms_per_step"].items()}, "pixel-only frac", round(d["roofline"]["frac"],3))
except Exception as e:
    print("$tag failed", e)
PY
}
source scripts/exp_list.sh
