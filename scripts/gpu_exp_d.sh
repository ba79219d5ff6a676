#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_configs.py -m gpu -q -x 2>&1 | tail -3
timeout 600 python scripts/fuzz_stress_gpu.py 6 31 2>&1 | tail -3
python scripts/stress_bench.py 16 2>&1 | tail -1
RMCV_WIDE_LABEL=0 python scripts/stress_bench.py 16 2>&1 | tail -1
for k in label_kernel order_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 4 -c 1 -f -o $OUT/stress_${k}_r02 python scripts/stress_bench.py 16 > $OUT/ncu_stress_$k.log 2>&1
  echo "ncu $k rc=$?"
done
