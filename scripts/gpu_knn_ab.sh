#!/bin/bash
# A/B of the self-k-NN kernels: GPU parity tests with the packet kernel, then the 1000-pair bench with variants.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short --timeout 600 -p no:cacheprovider -x > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -15 gpurun_out/pytest_gpu.log
for v in ${VARIANTS:-48}; do
  if [ "$v" = "0" ]; then export SB_KNN_PACKET=0; else export SB_KNN_PACKET=1 SB_KNN_PCAP=$v; fi
  timeout 600 python bench.py --frames 1000 --steps 3 --warmup 3 --no-e2e --cpu-seconds 0.1 > gpurun_out/knn_v_$v.log 2>&1
  echo "variant=$v exit $?"; python - <<PY
import json
for l in open("gpurun_out/knn_v_$v.log"):
    if l.startswith("{"):
        d = json.loads(l); print("variant=$v", d["ms_per_step"], d["roofline"]["stages_ms"], d["extras"].get("c3_knn_normals"))
PY
  if [ "$v" != "0" ]; then SB_KNN_STATS=1 timeout 600 python bench.py --frames 200 --steps 1 --warmup 3 --no-e2e --cpu-seconds 0.1 2>&1 | grep "self-knn" | sort | uniq -c | sort -rn | head -2; fi
done
