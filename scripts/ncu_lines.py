#!/usr/bin/env python
"""Per-CUDA-source-line profile of one kernel launch of an .ncu-rep captured with --import-source on:
   warp-stall samples, executed warp instructions and average active threads per line.
usage: ncu_lines.py rep kernel-regex [launch-index=0] [top=30]"""
import csv, io, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
idx = int(sys.argv[3]) if len(sys.argv) > 3 else 0
top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name",
                      "regex:" + kern, "--launch-skip", str(idx), "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
fname, hdr, lines = None, None, []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]; hdr = None
    elif r[0] == "Line No":
        hdr = r
    elif hdr and len(r) == len(hdr) and r[2] == "-":   # a source line (its SASS rows carry an address)
        g = lambda n: float(r[hdr.index(n)] or 0)
        lines.append((g("# Samples"), g("Instructions Executed"), g("Thread Instructions Executed"), fname, r[0], r[1].strip(),
                      g("stall_long_sb"), g("stall_wait"), g("stall_short_sb")))
ts, ti = sum(l[0] for l in lines), sum(l[1] for l in lines)
print(f"samples {ts:.0f}, warp instructions {ti:.0f}, active threads per instruction {sum(l[2] for l in lines) / max(ti, 1):.1f}")
print("  samples%   inst%  thr  long_sb%  file:line  source")
for l in sorted(lines, reverse=True)[:top]:
    print(f"  {100 * l[0] / ts:7.1f} {100 * l[1] / ti:7.1f} {l[2] / max(l[1], 1):4.0f} {100 * l[6] / max(l[0], 1):8.0f}  {l[3]}:{l[4]}  {l[5][:110]}")
