#!/bin/bash
for seg in 8 12 16 20 24 28 36 40; do echo -n "seg=$seg: "; RMCV_STRIP_SEG=$seg python scripts/bayer_bench_cold.py | tail -1; done
