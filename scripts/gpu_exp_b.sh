#!/bin/bash
# round-2 experiment set B: stage times of configs 4/5 and pipeline sensitivities of config 3
OUT=gpurun_out; mkdir -p $OUT
B="python bench.py --steps 10 --warmup 3 --no-cpu --no-extras --e2e-steps 1"
summ() { python - "$1" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print(sys.argv[1], "value", round(d["value"]), "ms/step", round(d["ms_per_step"],3), {k: round(v,3) for k,v in d["stage_ms_per_step"].items()}, "pix alone frac", round(d["roofline"]["frac"],3), "full", round(d["roofline"]["full_path_frac"],3))
except Exception as e: print(sys.argv[1], "failed", e)
PY
}
python scripts/stress_bench.py 16 2>&1 | tail -1
python scripts/stress_bench.py 64 2>&1 | tail -1
python scripts/latency_bench.py 2>&1 | tail -1
$B > $OUT/xb_base.json 2>/dev/null; summ $OUT/xb_base.json
RMCV_SERIAL=1 $B > $OUT/xb_serial.json 2>/dev/null; summ $OUT/xb_serial.json
RMCV_BGR_STRIP=1 $B > $OUT/xb_strip.json 2>/dev/null; summ $OUT/xb_strip.json
RMCV_BGR_STRIP=1 RMCV_SLOTS=5 $B --chunk 256 > $OUT/xb_strip_s5_c256.json 2>/dev/null; summ $OUT/xb_strip_s5_c256.json
RMCV_BGR_STRIP=1 $B --chunk 256 > $OUT/xb_strip_c256.json 2>/dev/null; summ $OUT/xb_strip_c256.json
RMCV_SLOTS=4 $B --chunk 296 > $OUT/xb_s4_c296.json 2>/dev/null; summ $OUT/xb_s4_c296.json
RMCV_PIX_S=3 RMCV_FRAME_RS=2048 $B > $OUT/xb_s3_rs2048.json 2>/dev/null; summ $OUT/xb_s3_rs2048.json
RMCV_PIX_S=2 RMCV_FRAME_RS=2048 $B > $OUT/xb_s2_rs2048.json 2>/dev/null; summ $OUT/xb_s2_rs2048.json
