#!/bin/bash
# small chunks + small pixel footprint + more slots: smooth co-residency of the labelling CTAs with the pixel CTAs
OUT=gpurun_out; mkdir -p $OUT
B="python bench.py --steps 10 --warmup 3 --no-cpu --no-extras --e2e-steps 1"
summ() { python - "$1" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print(sys.argv[1], "value", round(d["value"]), "ms/step", round(d["ms_per_step"],3), {k: round(v,3) for k,v in d["stage_ms_per_step"].items()}, "pix alone frac", round(d["roofline"]["frac"],3), "full", round(d["roofline"]["full_path_frac"],3))
except Exception as e: print(sys.argv[1], "failed", e)
PY
}
run() { tag=$1; shift; ch=$1; shift; env "$@" $B --chunk $ch > $OUT/xh_$tag.json 2>/dev/null; summ $OUT/xh_$tag.json; }
run c148_s5_rc2 148 RMCV_SLOTS=5 RMCV_PIX_RC=2
run c148_s8_rc2 148 RMCV_SLOTS=8 RMCV_PIX_RC=2
run c148_s5_rc4 148 RMCV_SLOTS=5
run c296_s5_rc2 296 RMCV_SLOTS=5 RMCV_PIX_RC=2
run c296_s4_rc2 296 RMCV_SLOTS=4 RMCV_PIX_RC=2
run c148_s6_rc2_prio1 148 RMCV_SLOTS=6 RMCV_PIX_RC=2 RMCV_PRIO=1
run c148_s6_rc2_prio2 148 RMCV_SLOTS=6 RMCV_PIX_RC=2 RMCV_PRIO=2
run c74_s8_rc2 74 RMCV_SLOTS=8 RMCV_PIX_RC=2
run c148_s6_rc2_rs2048 148 RMCV_SLOTS=6 RMCV_PIX_RC=2 RMCV_FRAME_RS=2048
run c222_s6_rc2 222 RMCV_SLOTS=6 RMCV_PIX_RC=2
