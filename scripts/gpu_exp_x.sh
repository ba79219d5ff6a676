#!/bin/bash
# Batch-1 latency anatomy: per-kernel durations in isolation (ncu launch list) and source-level stall samples of each kernel.
OUT=gpurun_out; mkdir -p $OUT
python scripts/latency_bench.py 3000 | tail -1
ncu --metrics gpu__time_duration.sum --clock-control none -s 120 -c 60 --csv --log-file $OUT/lat_launches.csv python scripts/latency_bench.py 60 > $OUT/lat_ncu.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/lat_launches.csv")) if len(r) > 10]
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value"); ui = hdr.index("Metric Unit")
d = collections.defaultdict(list)
for r in rows[1:]:
    v = float(r[vi].replace(",", "")); u = r[ui]
    d[r[ki][:60]].append(v / 1000 if u in ("ns", "nsecond") else v)
for k, v in d.items(): print("%-62s n=%d mean %.2f us" % (k, len(v), sum(v) / len(v)))
PY
for k in emit_kernel label_kernel contour_kernel fit_kernel order_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 30 -c 1 -f -o $OUT/lat_$k python scripts/latency_bench.py 60 > $OUT/lat_ncu_$k.log 2>&1
  echo "ncu $k rc=$?"
done
