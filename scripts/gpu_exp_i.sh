#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q -x > $OUT/pytest_gpu_i.log 2>&1; echo "pytest rc=$?"; tail -6 $OUT/pytest_gpu_i.log
B="python bench.py --steps 10 --warmup 3 --no-cpu --no-extras --e2e-steps 1"
summ() { python - "$1" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print(sys.argv[1], "value", round(d["value"]), "ms/step", round(d["ms_per_step"],3), {k: round(v,3) for k,v in d["stage_ms_per_step"].items()}, "pix alone frac", round(d["roofline"]["frac"],3), "full", round(d["roofline"]["full_path_frac"],3), "launches", d["gpu_launches"])
except Exception as e: print(sys.argv[1], "failed", e)
PY
}
$B > $OUT/xi_fused.json 2>/dev/null; summ $OUT/xi_fused.json
RMCV_FUSED_EMIT=0 $B > $OUT/xi_plain.json 2>/dev/null; summ $OUT/xi_plain.json
$B > $OUT/xi_fused2.json 2>/dev/null; summ $OUT/xi_fused2.json
RMCV_SERIAL=1 $B > $OUT/xi_fused_serial.json 2>/dev/null; summ $OUT/xi_fused_serial.json
python scripts/shim_timing.py 2>&1 | tail -1
