#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_configs.py -m gpu -q -x 2>&1 | tail -3
timeout 600 python scripts/fuzz_stress_gpu.py 6 31 2>&1 | tail -3
FUZZ_BIG=1 timeout 600 python scripts/fuzz_masks_gpu.py 6 32 2>&1 | tail -2
python scripts/stress_bench.py 16 2>&1 | tail -1
RMCV_WIDE_LABEL=0 python scripts/stress_bench.py 16 2>&1 | tail -1
python scripts/stress_bench.py 64 2>&1 | tail -1
RMCV_WIDE_LABEL=1 python scripts/stress_bench.py 64 2>&1 | tail -1
