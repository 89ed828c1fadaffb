#!/bin/bash
# Round-2 evidence in one gpurun call (one GPU): GPU tests, smoke, the default bench, the reference arm, then — only after
# the plain command has exited 0 without ncu — the per-launch metrics pass of one step and ncu --set full of the top kernels.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q --tb=short --timeout 900 -p no:cacheprovider -rs > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
tail -3 gpurun_out/smoke.log
timeout 1200 python bench.py > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/bench_default.err; echo "bench exit $?"
tail -c 300 gpurun_out/r02_bench_1gpu.json; echo
if [ "${SKIP_REF:-0}" != "1" ]; then
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference.json 2> gpurun_out/bench_reference.err; echo "ref exit $?"
tail -c 300 gpurun_out/r02_bench_reference.json; echo
fi
if [ "${NO_NCU:-0}" = "1" ]; then exit 0; fi
export SB_ICP_NOGRAPH=1
CMD="python bench.py --pairs 1024 --steps 1 --warmup 3 --no-e2e --no-sub --cpu-seconds 0.1"
$CMD > gpurun_out/plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum --clock-control none \
    -c 6000 --csv --log-file gpurun_out/r02_launches.csv $CMD > gpurun_out/ncu1.log 2>&1
echo "ncu metrics pass exit $?"
if [ "${NO_FULL:-0}" = "1" ]; then exit 0; fi
cap() {  # name regex skip count
  ncu --set full --clock-control none --import-source on -k regex:"$2" -s $3 -c $4 -f -o gpurun_out/r02_full_$1 $CMD > gpurun_out/ncu_$1.log 2>&1
  echo "ncu $1 exit $?"
}
cap self_knn "k_self_knn" 3 1
cap vox_insert "k_vox_insert" 3 1
cap normals "k_normals_from_graph|k_knn_redo" 6 2
cap icp_first "k_icp_match|k_icp_fallback|k_icp_accum|k_icp_solve" ${ICP_SKIP:-612} 8
