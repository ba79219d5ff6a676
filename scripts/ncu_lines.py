#!/usr/bin/env python
"""Per-source-line instruction and stall-sample totals from an ncu report with -lineinfo (--import-source on).
usage: ncu_lines.py REPORT.ncu-rep [top_n]"""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur_file = None; hdr = None; lines = []
for r in rows:
    if len(r) == 2 and r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if len(r) > 5 and r[0] == "Line No": hdr = r; continue
    if hdr and len(r) == len(hdr) and r[0].strip().isdigit():
        d = dict(zip(hdr[4:], r[4:]))
        def num(k):
            try: return float(d.get(k, "0").replace(",", ""))
            except ValueError: return 0.0
        lines.append((cur_file, int(r[0]), r[1].strip()[:110], num("Instructions Executed"), num("# Samples")))
tot_i = sum(l[3] for l in lines); tot_s = sum(l[4] for l in lines)
print(f"total warp instructions {tot_i:.0f}, samples {tot_s:.0f}")
print("--- by instructions")
for l in sorted(lines, key=lambda l: -l[3])[:top]:
    print(f"{100*l[3]/tot_i:5.1f}% inst {100*l[4]/max(tot_s,1):5.1f}% smp  {l[0]}:{l[1]}  {l[2]}")
print("--- by stall samples")
for l in sorted(lines, key=lambda l: -l[4])[:top]:
    print(f"{100*l[3]/tot_i:5.1f}% inst {100*l[4]/max(tot_s,1):5.1f}% smp  {l[0]}:{l[1]}  {l[2]}")
