#!/bin/bash
# Tail kernels for small chunks (label+contour / fit+order, or all four in one): parity and latency.
OUT=gpurun_out; mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
RMCV_TAIL=1 timeout 900 python -m pytest tests/test_gpu_detect.py tests/test_gpu_fuzz.py -m gpu -q -x 2>&1 | tail -2
for t in 2 1 0; do echo "RMCV_TAIL=$t"; RMCV_TAIL=$t python scripts/latency_bench.py 3000 | tail -1; done
python scripts/small_batch_bench.py | tail -1
RMCV_TAIL=0 python scripts/small_batch_bench.py | tail -1
