#!/bin/bash
# GPU parity tests, then the 1000-pair bench with and without the pair-resident tail kernel
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short --timeout 600 -p no:cacheprovider -x > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -15 gpurun_out/pytest_gpu.log
for v in ${VARIANTS:-64 0 16 256}; do
  SB_ICP_TAIL=$v timeout 600 python bench.py --frames 1000 --steps 3 --warmup 3 --no-e2e --cpu-seconds 0.1 > gpurun_out/icp_tail_$v.log 2>&1
  echo "tail=$v exit $?"; python - <<PY
import json
for l in open("gpurun_out/icp_tail_$v.log"):
    if l.startswith("{"):
        d = json.loads(l); print("tail=$v", d["ms_per_step"], d["roofline"]["stages_ms"], {k: d["extras"].get("c2_streaming", {}).get(k) for k in ("mean_ms", "p50_ms", "p99_ms", "ms_per_frame_mean")}, d["extras"].get("error"))
PY
done
