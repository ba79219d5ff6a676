#!/bin/bash
python scripts/bayer_bench_cold.py | tail -1
python scripts/bayer_detect_bench.py | tail -1
RMCV_STRIP_SEG=60 python scripts/bayer_detect_bench.py | tail -1
RMCV_STRIP_SEG=32 python scripts/bayer_detect_bench.py | tail -1
RMCV_STRIP_SEG=16 python scripts/bayer_detect_bench.py | tail -1
timeout 600 python -m pytest tests/test_gpu_bayer_strip.py tests/test_gpu_pixel.py -m gpu -q -x 2>&1 | tail -2
