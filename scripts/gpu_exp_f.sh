#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_legacy.py tests/test_gpu_shim.py -m gpu -q -x 2>&1 | tail -8
timeout 600 python scripts/fuzz_legacy_gpu.py 40 11 2>&1 | tail -4
