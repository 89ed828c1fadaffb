#!/bin/bash
# ncu --set full of the top kernels of one C5 step (1024 pairs), one launch each, source imported
mkdir -p gpurun_out
export SB_ICP_NOGRAPH=1
CMD="python bench.py --pairs 1024 --steps 1 --warmup 3 --no-e2e --no-sub --cpu-seconds 0.1"
$CMD > gpurun_out/plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain.log; exit 1; }
cap() {  # name regex skip count
  ncu --set full --clock-control none --import-source on -k regex:"$2" -s $3 -c $4 -f -o gpurun_out/r02_full_$1 $CMD > gpurun_out/ncu_$1.log 2>&1
  echo "ncu $1 exit $?"
}
cap self_knn "k_self_knn" 3 1
cap vox_insert "k_vox_insert" 3 1
cap normals "k_normals_from_graph|k_knn_redo" 6 2
cap icp_first "k_icp_match|k_icp_fallback|k_icp_accum|k_icp_solve" ${ICP_SKIP:-612} 8
