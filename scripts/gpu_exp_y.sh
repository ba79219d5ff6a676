#!/bin/bash
# Chained launches (programmatic dependent launch) for small chunks: parity, latency, gap anatomy, small-batch throughput.
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_detect.py tests/test_gpu_fuzz.py tests/test_gpu_shim.py -m gpu -q -x 2>&1 | tail -3
python scripts/latency_bench.py 3000 | tail -1
RMCV_CHAINED=0 python scripts/latency_bench.py 3000 | tail -1
python scripts/phase_stamps.py 2>&1 | head -4
python scripts/small_batch_bench.py | tail -1
RMCV_CHAINED=0 python scripts/small_batch_bench.py | tail -1
