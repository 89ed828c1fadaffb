#!/bin/bash
# GPU parity tests + a reduced C5 bench (512 pairs) with sub-results
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short --timeout 600 -p no:cacheprovider -x > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -15 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --pairs ${PAIRS:-1024} --steps 3 --warmup 3 ${BENCH_ARGS} > gpurun_out/bench_c5_small.log 2>&1; echo "bench exit $?"
tail -c 3000 gpurun_out/bench_c5_small.log | grep -v "^{" | tail -20
python - <<PY
import json
for l in open("gpurun_out/bench_c5_small.log"):
    if l.startswith("{"):
        d = json.loads(l)
        print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"] and d["e2e"]["value"], "launches", d["gpu_launches"])
        print("stages", d["roofline"]["stages_ms"]); print("stats", d["workload_stats"]); print("cpu", d["cpu_baseline"])
        for k in ("c2_batch", "c2_streaming", "c3_knn_normals", "c4_loop_closure", "error"):
            if k in d: print(k, {a: b for a, b in d[k].items() if a not in ("slowest_frames",)} if isinstance(d[k], dict) else d[k])
PY
