#!/bin/bash
# co-residency experiment: cap the labelling kernels' CTAs per SM (shared-memory padding) so that pixel CTAs keep their slots
OUT=gpurun_out; mkdir -p $OUT
B="python bench.py --steps 10 --warmup 3 --no-cpu --no-extras --e2e-steps 1"
summ() { python - "$1" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print(sys.argv[1], "value", round(d["value"]), "ms/step", round(d["ms_per_step"],3), {k: round(v,3) for k,v in d["stage_ms_per_step"].items()}, "full", round(d["roofline"]["full_path_frac"],3))
except Exception as e: print(sys.argv[1], "failed", e)
PY
}
run() { tag=$1; shift; env "$@" $B > $OUT/xj_$tag.json 2>/dev/null; summ $OUT/xj_$tag.json; }
run base RMCV_SLOTS=3
run rc2_pad30 RMCV_PIX_RC=2 RMCV_CHAIN_PAD=30000
run rc2_pad45 RMCV_PIX_RC=2 RMCV_CHAIN_PAD=45000
run rc2_pad60 RMCV_PIX_RC=2 RMCV_CHAIN_PAD=60000
run rc2_pad90 RMCV_PIX_RC=2 RMCV_CHAIN_PAD=90000
run rc4_pad36 RMCV_CHAIN_PAD=36000
run rc4_pad70 RMCV_CHAIN_PAD=70000
run rc2_pad45_s5 RMCV_PIX_RC=2 RMCV_CHAIN_PAD=45000 RMCV_SLOTS=5
run rc2_pad45_fused RMCV_PIX_RC=2 RMCV_CHAIN_PAD=45000 RMCV_FUSED_EMIT=1
run rc2_pad45_prio1 RMCV_PIX_RC=2 RMCV_CHAIN_PAD=45000 RMCV_PRIO=1
timeout 300 python scripts/fused_emit_parity_gpu.py 2>&1 | tail -1
python scripts/shim_timing.py 2>&1 | tail -1
