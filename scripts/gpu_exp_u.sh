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
run() { tag=$1; shift; env "$@" $B > $OUT/xu_$tag.json 2>/dev/null; summ $OUT/xu_$tag.json; }
run base A=1
run base2 A=1
run fused RMCV_FUSED_EMIT=1
timeout 900 python -m pytest tests/test_gpu_detect.py tests/test_gpu_configs.py tests/test_gpu_golden.py -m gpu -q -x 2>&1 | tail -2
python scripts/stress_bench.py 16 | tail -1
python scripts/latency_bench.py | tail -1
python scripts/bayer_detect_bench.py | tail -1
