#!/usr/bin/env python
"""SASS-level execution profile of one kernel of an .ncu-rep: opcode mix and segments of equal execution count."""
import collections, csv, io, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "rows": []}
        blocks.append(cur)
    elif cur is not None and cur["hdr"] is None and r and r[0] == "Address":
        cur["hdr"] = r
    elif cur is not None and cur["hdr"] is not None and len(r) == len(cur["hdr"]):
        cur["rows"].append(r)
b = blocks[int(sys.argv[3]) if len(sys.argv) > 3 else 0]
hdr = b["hdr"]
isrc, iex, ism, ith = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Avg. Threads Executed")
data = [(r[isrc].strip(), int(r[iex]), int(r[ism]), float(r[ith] or 0)) for r in b["rows"]]
tot, ts = sum(d[1] for d in data), sum(d[2] for d in data)
print(b["name"][:100]); print("total warp-inst", tot, "samples", ts, "sass lines", len(data))
op = collections.Counter(); ops = collections.Counter()
for s, e, sm, t in data:
    k = s.split()[1] if s.startswith("@") else s.split()[0]
    op[k] += e; ops[k] += sm
for k, v in op.most_common(22):
    print(f"{k:16s} {v:12d} {100 * v / tot:5.1f}%  samples {100 * ops[k] / max(ts, 1):5.1f}%")
print("--- segments of equal execution count: [sass idx range] exec, threads, samples ---")
start, prev = 0, None
for i, d in enumerate(data + [("", -1, 0, 0)]):
    if prev is None:
        prev = d[1]; continue
    if d[1] < 0 or abs(d[1] - prev) > 0.25 * max(prev, 1):
        seg = data[start:i]
        print(f"[{start:4d}-{i - 1:4d}] n={i - start:4d} exec={prev:11d} ({100 * sum(x[1] for x in seg) / tot:5.1f}% of inst) thr={seg[0][3]:4.0f} samples={100 * sum(x[2] for x in seg) / max(ts, 1):5.1f}%  {seg[0][0][:50]}")
        start, prev = i, d[1]
