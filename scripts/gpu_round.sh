#!/bin/bash
# Round evidence in one gpurun call: GPU tests, smoke, the default bench, the launch list of one step and one
# `ncu --set full` capture of the top kernels (host-driven ICP loop so that every kernel is an ordinary launch).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q --tb=short --timeout 600 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
tail -2 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench_default.log 2>&1; echo "bench exit $?" >> gpurun_out/bench_default.log
tail -c 400 gpurun_out/bench_default.log
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.log 2>&1; echo "ref exit $?" >> gpurun_out/bench_reference.log
tail -c 600 gpurun_out/bench_reference.log
export SB_ICP_NOGRAPH=1
CMD="python bench.py --frames 1000 --steps 1 --warmup 3 --no-e2e --cpu-seconds 0.1"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu1.log 2>&1
echo "ncu launch list exit $?"
$CMD > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"k_knn|k_icp_match|k_icp_fallback|k_vox_insert|k_icp_accum" -s ${NCU_SKIP:-15} -c ${NCU_COUNT:-8} -f -o gpurun_out/prof_round $CMD > gpurun_out/ncu2.log 2>&1
echo "ncu full exit $?"
