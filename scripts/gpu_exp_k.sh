#!/bin/bash
python scripts/bayer_bench_cold.py | tail -1
for seg in 24 32 48 64 96; do for mb in 3 4; do echo -n "seg=$seg minb=$mb: "; RMCV_STRIP_SEG=$seg RMCV_STRIP_MINB=$mb python scripts/bayer_bench_cold.py | tail -1; done; done
echo -n "generic: "; RMCV_BAYER_GENERIC=1 python scripts/bayer_bench_cold.py | tail -1
