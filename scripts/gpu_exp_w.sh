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
run() { tag=$1; shift; env "$@" $B > $OUT/xw_$tag.json 2>/dev/null; summ $OUT/xw_$tag.json; }
run warm A=1
run base A=1
run fused RMCV_FUSED_EMIT=1
run gy1 RMCV_CONTOUR_GY=1
run gy8 RMCV_CONTOUR_GY=8
run rs4096 RMCV_FRAME_RS=4096
run s2 RMCV_SLOTS=2
run base_b A=1
