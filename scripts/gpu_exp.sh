#!/bin/bash
# experiment runner: each "run TAG [bench args --] ENV..." runs the short bench with that environment
OUT=gpurun_out; mkdir -p $OUT
run() { tag=$1; shift; extra=""; if [ "$1" = "--args" ]; then extra="$2"; shift 2; fi
  env "$@" python bench.py --steps 5 --warmup 3 --batch 1024 --no-cpu --no-extras --e2e-steps 1 $extra > $OUT/exp_$tag.json 2> $OUT/exp_$tag.err
  python - <<PY
import json
try:
    d=json.load(open("$OUT/exp_$tag.json"))
    print("$tag: value", round(d["value"]), "ms/step", round(d["ms_per_step"],3), "stage", {k:round(v,3) for k,v in d["stage_ms_per_step"].items()}, "chunk", d["config"]["chunk_frames"], "e2e", round(d["e2e"]["value"]))
except Exception as e:
    print("$tag failed", e)
PY
}
source scripts/exp_list.sh
