#!/bin/bash
for cfg in "64 4" "32 4" "16 8" "64 8" "128 2" "32 16"; do set -- $cfg
SB_VOX_GRID=$1 SB_VOX_TPC=$2 timeout 600 python bench.py --pairs 1024 --steps 3 --warmup 3 --no-e2e --no-sub --cpu-seconds 0.1 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('grid=$1 tpc=$2 voxel', round(d['roofline']['stages_ms']['voxel'],3), 'step', round(d['ms_per_step'],2))"
done
