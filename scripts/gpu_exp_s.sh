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
run() { tag=$1; shift; env "$@" $B > $OUT/xs_$tag.json 2>/dev/null; summ $OUT/xs_$tag.json; }
run base A=1
run gy4 RMCV_CONTOUR_GY=4
run gy2 RMCV_CONTOUR_GY=2
run gy16 RMCV_CONTOUR_GY=16
run rs2048 RMCV_FRAME_RS=2048
run rs4096 RMCV_FRAME_RS=4096
run emitbh16 RMCV_EMIT_BH=16
run emitbh64 RMCV_EMIT_BH=64
run prio1 RMCV_PRIO=1
run labstreams1 RMCV_LAB_STREAMS=1
run strip RMCV_BGR_STRIP=1
