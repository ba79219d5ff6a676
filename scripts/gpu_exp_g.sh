#!/bin/bash
# pipeline sensitivity to the pixel kernel's shared-memory footprint (generic-geometry kernel)
OUT=gpurun_out; mkdir -p $OUT
B="python bench.py --steps 10 --warmup 3 --no-cpu --no-extras --e2e-steps 1"
summ() { python - "$1" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print(sys.argv[1], "value", round(d["value"]), "ms/step", round(d["ms_per_step"],3), {k: round(v,3) for k,v in d["stage_ms_per_step"].items()}, "pix alone frac", round(d["roofline"]["frac"],3), "full", round(d["roofline"]["full_path_frac"],3))
except Exception as e: print(sys.argv[1], "failed", e)
PY
}
run() { tag=$1; shift; env "$@" $B > $OUT/xg_$tag.json 2>/dev/null; summ $OUT/xg_$tag.json; }
run generic RMCV_PIX_GENERIC=1
run rc2 RMCV_PIX_RC=2
run rc2_nt160 RMCV_PIX_RC=2 RMCV_PIX_NT=160
run rc2_s6 RMCV_PIX_RC=2 RMCV_PIX_S=6
run rc2_s3 RMCV_PIX_RC=2 RMCV_PIX_S=3
run rc3 RMCV_PIX_RC=3
run rc2_rs2048 RMCV_PIX_RC=2 RMCV_FRAME_RS=2048
run rc1_s8 RMCV_PIX_RC=1 RMCV_PIX_S=8 RMCV_PIX_NT=128
run rc2_bh16 RMCV_PIX_RC=2 RMCV_PIX_BH=16
run rc2_bh64 RMCV_PIX_RC=2 RMCV_PIX_BH=64
