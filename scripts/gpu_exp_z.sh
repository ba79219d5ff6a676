#!/bin/bash
# Label-kernel phase work (asynchronous pointer jumping, Euler terms in shared memory, nested verdicts in the label kernel):
# parity, latency, anatomy, headline step.
OUT=gpurun_out; mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
python scripts/latency_bench.py 3000 | tail -1
python scripts/phase_stamps.py 2>&1 | head -9
python bench.py --steps 20 --warmup 3 --no-cpu --no-extras --e2e-steps 1 > $OUT/xz_base.json 2> $OUT/xz_base.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/xz_base.json"))
print("value", round(d["value"]), "ms/step", round(d["ms_per_step"], 3), {k: round(v, 3) for k, v in d["stage_ms_per_step"].items()}, "full", round(d["roofline"]["full_path_frac"], 3))
PY
python scripts/stress_bench.py 16 | tail -1
