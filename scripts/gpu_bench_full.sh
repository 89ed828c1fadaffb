#!/bin/bash
mkdir -p gpurun_out
nproc; free -g | head -2
( time timeout 1200 python bench.py > gpurun_out/bench_default.log 2>&1 ) 2>&1 | grep real; echo "bench exit $?"
grep -v "^{" gpurun_out/bench_default.log | tail -5
python - <<PY
import json
for l in open("gpurun_out/bench_default.log"):
    if l.startswith("{"):
        d = json.loads(l)
        print("value", d["value"], "ms", d["ms_per_step"], "launches", d["gpu_launches"], "clocks", d["clocks"])
        print("e2e", d["e2e"]); print("stages", d["roofline"]["stages_ms"]); print("roofline", {k: v for k, v in d["roofline"].items() if k not in ("stages_ms",)})
        print("stats", d["workload_stats"]); print("cpu", {k: v for k, v in d["cpu_baseline"].items() if k != "sample"})
        for k in ("c2_batch", "c2_streaming", "c3_knn_normals", "c4_loop_closure", "error"):
            if k in d: print(k, {a: b for a, b in d[k].items() if a not in ("slowest_frames", "workload")} if isinstance(d[k], dict) else d[k])
PY
( time timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.log 2>&1 ) 2>&1 | grep real
tail -c 1500 gpurun_out/bench_reference.log
