#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total and share."""
import collections
import csv
import sys

path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/launches.csv"
last = int(sys.argv[2]) if len(sys.argv) > 2 else 0   # only the last N launches (one step)
rows = [r for r in csv.reader(open(path)) if len(r) > 10]
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
seq = []
for r in rows[1:]:
    try:
        seq.append((r[ki].split("(")[0].replace("void ", ""), float(r[vi].replace(",", ""))))
    except ValueError:
        pass
if last:
    seq = seq[-last:]
agg = collections.OrderedDict()
for n, v in seq:
    agg.setdefault(n, []).append(v)
tot = sum(sum(v) for v in agg.values())
print(f"{len(seq)} launches, {tot / 1e6:.3f} ms of kernel time")
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print(f"{k:34s} n={len(v):5d} sum={sum(v) / 1e3:11.1f} us {100 * sum(v) / tot:5.1f}%  max={max(v) / 1e3:9.1f} us")
if "--seq" in sys.argv:
    name = sys.argv[sys.argv.index("--seq") + 1]
    print(name, [round(v / 1e3) for n, v in seq if n == name])
