#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short --timeout 600 -p no:cacheprovider -x -k "voxel or float32 or smoke or register" > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
for v in ${VARIANTS:-48}; do
  SB_VOX_WINDOW=$v timeout 600 python bench.py --pairs 1024 --steps 3 --warmup 3 --no-e2e --no-sub --cpu-seconds 0.1 > gpurun_out/vox_$v.log 2>&1
  echo "window=$v exit $?"; python - <<PY
import json
for l in open("gpurun_out/vox_$v.log"):
    if l.startswith("{"):
        d = json.loads(l); print("window=$v", d["ms_per_step"], d["roofline"]["stages_ms"]["voxel"])
PY
done
