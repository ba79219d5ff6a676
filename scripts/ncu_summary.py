#!/usr/bin/env python
"""Condense ncu output into the small text files kept under profiles/.

    python scripts/ncu_summary.py launches gpurun_out/launches_TAG.csv  > profiles/rNN_launches_TAG.md
    python scripts/ncu_summary.py full     gpurun_out/KERNEL_TAG.ncu-rep > profiles/rNN_KERNEL_TAG.md

`launches`: per-kernel count / total / mean of gpu__time_duration.sum and each kernel's share of the GPU time.
`full`: the handful of `--set full` metrics the roofline entry of bench.py quotes (DRAM bytes, throughput, occupancy).
"""
import csv
import subprocess
import sys
from collections import OrderedDict

FULL_METRICS = [
    "gpu__time_duration.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__bytes_read.sum.per_second", "dram__bytes_write.sum.per_second",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__cycles_elapsed.max",
    "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "launch__occupancy_limit_warps",
    "launch__waves_per_multiprocessor",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__inst_executed_op_shared_atom.sum", "l1tex__t_set_accesses_pipe_lsu_mem_global_op_atom.sum",
    "lts__t_sectors_op_atom.sum", "lts__t_sectors_op_red.sum",
    "smsp__average_warp_latency_issue_stalled_barrier.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
]


def launches(path):
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr = rows[0]
    ik, iv, ig, ib = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size"), hdr.index("Block Size")
    agg = OrderedDict()
    for r in rows[1:]:
        key = (r[ik], r[ig], r[ib])
        a = agg.setdefault(key, [0, 0.0])
        a[0] += 1
        a[1] += float(r[iv].replace(",", ""))
    total = sum(a[1] for a in agg.values())
    print(f"# ncu launch list: {path}\n")
    print("gpu__time_duration.sum per launch, `--clock-control none`; serialised, cold-cache times — use the SHARES.\n")
    print("| kernel | grid | block | launches | total us | mean us | share |")
    print("|---|---|---|---:|---:|---:|---:|")
    for (k, g, b), (n, ns) in agg.items():
        print(f"| `{k}` | {g} | {b} | {n} | {ns / 1e3:.1f} | {ns / 1e3 / n:.2f} | {100 * ns / total:.1f}% |")
    print(f"\ntotal {total / 1e3:.1f} us over {sum(a[0] for a in agg.values())} launches")


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    print(f"# ncu --set full: {path}\n")
    for n, r in enumerate(rows[2:]):
        get = lambda name: r[hdr.index(name)] if name in hdr else None
        print(f"## launch {n}: `{get('Kernel Name')}` grid {get('Grid Size')} block {get('Block Size')}\n")
        print("| metric | value | unit |")
        print("|---|---:|---|")
        for m in FULL_METRICS:
            if m in hdr:
                print(f"| {m} | {r[hdr.index(m)]} | {units[hdr.index(m)]} |")
        rd, wr = get("dram__bytes_read.sum"), get("dram__bytes_write.sum")
        if rd and wr:
            ur, uw = units[hdr.index("dram__bytes_read.sum")], units[hdr.index("dram__bytes_write.sum")]
            scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            tot = float(rd.replace(",", "")) * scale.get(ur, 1.0) + float(wr.replace(",", "")) * scale.get(uw, 1.0)
            print(f"| **dram traffic (read+write)** | {tot:.0f} | byte |")
        print()


def traffic(path, frames_per_launch):
    """profiles/pixel_traffic.json for bench.py: DRAM bytes per frame of the pixel kernel from one `--set full` capture."""
    import json
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, r = rows[0], rows[1], rows[2]
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    def val(name):
        i = hdr.index(name)
        return float(r[i].replace(",", "")) * scale.get(units[i], 1.0)
    rd, wr = val("dram__bytes_read.sum"), val("dram__bytes_write.sum")
    grid = r[hdr.index("Grid Size")]
    print(json.dumps({"source": path, "kernel": r[hdr.index("Kernel Name")], "grid": grid, "frames_per_launch": int(frames_per_launch),
                      "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes_per_frame": (rd + wr) / float(frames_per_launch),
                      "gpu_time_us": float(r[hdr.index("gpu__time_duration.sum")].replace(",", ""))}, indent=1))


if __name__ == "__main__":
    if sys.argv[1] == "traffic":
        traffic(sys.argv[2], sys.argv[3])
    else:
        {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
