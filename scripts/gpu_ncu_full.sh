#!/bin/bash
# One `ncu --set full` capture of selected kernels of a bench step (host-driven ICP loop so every kernel is a launch).
# usage: KERNELS='regex' SKIP=n COUNT=n FRAMES=n OUT=name bash scripts/gpu_ncu_full.sh
mkdir -p gpurun_out
export SB_ICP_NOGRAPH=1
CMD="python bench.py --frames ${FRAMES:-250} --steps 1 --warmup 3 --no-e2e --cpu-seconds 0.1"
$CMD > gpurun_out/plain_full.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"${KERNELS:-k_vox_insert|k_knn}" -s ${SKIP:-6} -c ${COUNT:-2} -f -o gpurun_out/${OUT:-prof} $CMD > gpurun_out/ncu_full.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/ncu_full.log | cut -c1-300
