#!/bin/bash
# per-pass kernel times of the ICP loop (host-driven loop so that every pass is an ordinary launch)
mkdir -p gpurun_out
export SB_ICP_NOGRAPH=1
CMD="python bench.py --pairs 1024 --steps 1 --warmup 3 --no-e2e --no-sub --cpu-seconds 0.1"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu1.log 2>&1
echo "ncu launch list exit $?"
python scripts/launch_summary.py gpurun_out/launches.csv 0 --seq k_icp_match | head -30
python scripts/launch_summary.py gpurun_out/launches.csv 0 --seq k_icp_match | tail -1 | tr ',' '\n' | tail -60 | tr '\n' ' '
SB_ICP_STATS=1 python bench.py --pairs 1024 --steps 1 --warmup 3 --no-e2e --no-sub --cpu-seconds 0.1 2>&1 | grep "icp iterations"
