#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_shim.py tests/test_gpu_multi.py tests/test_gpu_golden.py -m gpu -q -x 2>&1 | tail -4
timeout 600 python bench.py --no-cpu > $OUT/bench_o.json 2> $OUT/bench_o.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open("$OUT/bench_o.json"))
print("value", round(d["value"]), "ms/step", round(d["ms_per_step"],3), "full", round(d["roofline"]["full_path_frac"],3), "e2e", round(d["e2e"]["value"]))
x=d["extras"]
print("latency", x["latency_batch1"]["p50_us"], x["latency_batch1"]["p99_us"])
print("bayer pix", x["bayer_pixel_stage"]["frac_of_peak"], "bayer detect", x["bayer_full_detect"]["frames_per_s"], x["bayer_full_detect"]["full_path_frac_of_hbm_peak"], x["bayer_full_detect"].get("e2e_from_host_frames_per_s"))
print("stress", {k:(round(v["frames_per_s"]), round(v["full_path_frac_of_hbm_peak"],3)) for k,v in x["stress_4096x3072"].items() if k.startswith("batch")})
PY
