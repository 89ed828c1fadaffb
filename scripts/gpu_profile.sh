#!/bin/bash
# Full-size bench, then ncu launch list + one full capture of the top kernels (host-driven ICP loop so that every
# kernel is an ordinary launch under the profiler).
mkdir -p gpurun_out
timeout 1200 python bench.py > gpurun_out/bench_full.log 2>&1; echo "bench exit $?" >> gpurun_out/bench_full.log
tail -3 gpurun_out/bench_full.log
export SB_ICP_NOGRAPH=1
CMD="python bench.py --frames ${NCU_FRAMES:-24} --steps 1 --warmup 3 --no-e2e --cpu-seconds 0.1"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu1.log 2>&1
echo "ncu1 exit $?"
$CMD > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"k_icp_iter|k_knn|k_sort_scatter" -s 12 -c 6 -o gpurun_out/prof_r01 $CMD > gpurun_out/ncu2.log 2>&1
echo "ncu2 exit $?"
ls -la gpurun_out
