#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
B="python bench.py --steps 20 --warmup 3 --no-cpu --no-extras --e2e-steps 1"
summ() { python - "$1" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print(sys.argv[1], "value", round(d["value"]), "ms/step", round(d["ms_per_step"],3), {k: round(v,3) for k,v in d["stage_ms_per_step"].items()}, "full", round(d["roofline"]["full_path_frac"],3))
except Exception as e: print(sys.argv[1], "failed", e)
PY
}
run() { tag=$1; shift; env "$@" $B > $OUT/xv_$tag.json 2>/dev/null; summ $OUT/xv_$tag.json; }
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
run flags A=1
run flags2 A=1
run serial RMCV_SERIAL=1
