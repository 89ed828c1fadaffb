#!/bin/bash
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x -k "voxel or float32 or register" 2>&1 | tail -2
timeout 600 python bench.py --pairs 1024 --steps 3 --warmup 3 --no-e2e --no-sub --cpu-seconds 0.1 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('step', round(d['ms_per_step'],2), {k: round(v,3) for k,v in d['roofline']['stages_ms'].items()})"
