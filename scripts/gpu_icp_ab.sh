#!/bin/bash
# ICP loop variants: open lanes of an item from which the packet traversal takes over
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short --timeout 600 -p no:cacheprovider -x > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
for v in ${VARIANTS:-6}; do
  echo "packet_min=$v"; SB_ICP_PACKET_MIN=$v timeout 600 python scripts/icp_tail_cost.py 2>&1 | tail -7
done
timeout 600 python scripts/stream_latency.py --frames 150 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print({k:d[k] for k in ('ms_per_frame_mean','p50','p99','max','split_ms_mean','launches_per_frame')})"
