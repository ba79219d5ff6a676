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
run() { tag=$1; shift; ch=$1; shift; env "$@" $B --chunk $ch > $OUT/xq_$tag.json 2>/dev/null; summ $OUT/xq_$tag.json; }
run c1024 1024 A=1
run c1024_fused 1024 RMCV_FUSED_EMIT=1
run c768 768 A=1
run c896 896 A=1
run c1024_s2 1024 RMCV_SLOTS=2
run c1024_s4 1024 RMCV_SLOTS=4
run c1024_b 1024 A=1
env A=1 $B --batch 2048 --chunk 1024 > $OUT/xq_b2048_c1024.json 2>/dev/null; summ $OUT/xq_b2048_c1024.json
env A=1 $B --batch 2048 --chunk 592 > $OUT/xq_b2048_c592.json 2>/dev/null; summ $OUT/xq_b2048_c592.json
env A=1 $B --batch 2048 --chunk 2048 > $OUT/xq_b2048_c2048.json 2>/dev/null; summ $OUT/xq_b2048_c2048.json
