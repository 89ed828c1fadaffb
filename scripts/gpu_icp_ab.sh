#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short --timeout 600 -p no:cacheprovider -x > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
timeout 600 python scripts/icp_tail_cost.py 2>&1 | tail -7
SB_ICP_STATS=1 timeout 600 python bench.py --pairs 1024 --steps 3 --warmup 3 --no-e2e --no-sub --cpu-seconds 0.1 2>&1 | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['ms_per_step'], d['roofline']['stages_ms'])
    elif 'icp iterations' in l: print(l.strip())"
