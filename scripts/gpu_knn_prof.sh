#!/bin/bash
# stats of the packet self-k-NN kernel + one ncu --set full capture of it
mkdir -p gpurun_out
SB_KNN_STATS=1 timeout 600 python bench.py --frames 300 --steps 1 --warmup 3 --no-e2e --cpu-seconds 0.1 > gpurun_out/knn_stats.log 2>&1
grep "self-knn" gpurun_out/knn_stats.log | sort | uniq -c | sort -rn | head -5
CMD="python bench.py --frames 500 --steps 1 --warmup 3 --no-e2e --cpu-seconds 0.1"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_self_knn|k_knn_redo|k_normals_from" -s 9 -c 3 -f -o gpurun_out/knn_packet $CMD > gpurun_out/ncu_knn.log 2>&1
echo "ncu exit $?"
