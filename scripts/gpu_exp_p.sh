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
run() { tag=$1; shift; ch=$1; shift; env "$@" $B --chunk $ch > $OUT/xp_$tag.json 2>/dev/null; summ $OUT/xp_$tag.json; }
run c592 592 A=1
run c512 512 A=1
run c1024 1024 A=1
run c342 342 A=1
run c512_s4 512 RMCV_SLOTS=4
run c444 444 A=1
run c592_fused 592 RMCV_FUSED_EMIT=1
run c512_fused 512 RMCV_FUSED_EMIT=1
python __graft_entry__.py --smoke 2>&1 | tail -1
