#!/bin/bash
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x 2>&1 | tail -2
for v in 16 32; do
SB_KNN_TOT10=$v SB_KNN_STATS=1 python - 2>&1 <<PY | grep -v "^\[slam_b200\] self-knn k=10.*" | tail -3
import sys, os
sys.path.insert(0, "."); sys.path.insert(0, "tests"); sys.path.insert(0, "lidar-slam-from-scratch_b200/python")
import torch, bench, oracle_lib, slam_b200
eng = slam_b200.Engine(0); syn = oracle_lib.Synth()
r = bench.sub_c3(eng, syn, torch)
print("tot10=$v", r["ms"], r["knn_plus_normals_queries_per_s"])
PY
done
SB_KNN_TOT10=16 SB_KNN_STATS=1 python bench.py --pairs 256 --steps 1 --warmup 3 --no-e2e --cpu-seconds 0.1 2>&1 | grep "self-knn" | sort | uniq -c | sort -rn | head -3
for v in 16 32; do
SB_KNN_TOT10=$v python - 2>&1 <<PY | tail -1
import sys, os
sys.path.insert(0, "."); sys.path.insert(0, "tests"); sys.path.insert(0, "lidar-slam-from-scratch_b200/python")
import torch, bench, oracle_lib, slam_b200
eng = slam_b200.Engine(0); syn = oracle_lib.Synth()
r = bench.sub_c3(eng, syn, torch)
print("tot10=$v (no stats)", r["ms"], r["knn_plus_normals_queries_per_s"])
PY
done
python bench.py --pairs 1024 --steps 3 --warmup 3 --no-e2e --no-sub --cpu-seconds 0.1 2>&1 | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['ms_per_step'], d['roofline']['stages_ms'])"
