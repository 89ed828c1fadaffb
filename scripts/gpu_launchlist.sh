#!/bin/bash
# Launch list (device time of every kernel) of one full-size bench step with the host-driven ICP loop.
mkdir -p gpurun_out
export SB_ICP_NOGRAPH=1
CMD="python bench.py --frames ${NCU_FRAMES:-1000} --steps 1 --warmup 3 --no-e2e --cpu-seconds 0.1"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s ${NCU_SKIP:-0} -c ${NCU_COUNT:-1500} --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu1.log 2>&1
echo "ncu1 exit $?"
tail -c 600 gpurun_out/plain.log
