#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short --timeout 600 -p no:cacheprovider -x > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
for v in ${VARIANTS:-256 128 512}; do
for ppc in ${PPC:-4096}; do
  SB_ICP_SUB=$v timeout 900 python bench.py --steps 3 --warmup 3 --no-sub --cpu-seconds 0.1 --pairs-per-call $ppc > gpurun_out/e2e_${v}_$ppc.log 2>&1
  echo "sub=$v ppc=$ppc exit $?"; grep -v "^{" gpurun_out/e2e_${v}_$ppc.log | tail -3; python - <<PY
import json
for l in open("gpurun_out/e2e_${v}_$ppc.log"):
    if l.startswith("{"):
        d = json.loads(l); print("sub=$v ppc=$ppc value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "GB/s", d["e2e"]["h2d_bytes_per_step"] / d["e2e"]["ms_per_step"] / 1e6)
PY
done; done
