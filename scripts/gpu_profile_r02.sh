#!/bin/bash
# Round-2 evidence in one gpurun call (after the plain command has exited 0 without ncu):
#  A. per-launch time / DRAM bytes / warp instructions of one whole step (host-driven ICP loop: every pass an ordinary launch)
#  B. ncu --set full of the top kernels (source imported)
mkdir -p gpurun_out
export SB_ICP_NOGRAPH=1
CMD="python bench.py --pairs 1024 --steps 1 --warmup 3 --no-e2e --no-sub --cpu-seconds 0.1"
$CMD > gpurun_out/plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum --clock-control none \
    -c 6000 --csv --log-file gpurun_out/r02_launches.csv $CMD > gpurun_out/ncu1.log 2>&1
echo "ncu metrics pass exit $?"
$CMD > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"k_self_knn|k_icp_match|k_icp_fallback|k_vox_insert|k_normals_from_graph|k_knn_redo|k_icp_accum" \
    -s ${NCU_SKIP:-21} -c ${NCU_COUNT:-14} -f -o gpurun_out/r02_full $CMD > gpurun_out/ncu2.log 2>&1
echo "ncu full exit $?"
