#!/bin/bash
# A/B/C... of several builds of the library in one gpurun call: bench.py (device-resident C5 step only) on every
# lidar-slam-from-scratch_b200/libslam_b200_<name>.so named in LIBS (or the current build with the environment
# ENV_<name>), ROUNDS times in alternation.
mkdir -p gpurun_out
P=lidar-slam-from-scratch_b200
for r in $(seq 1 ${ROUNDS:-2}); do
for v in $LIBS; do
  lib=$P/libslam_b200_$v.so
  [ -f $lib ] || lib=$P/libslam_b200.so      # a name without a build of its own: the current build + ENV_<name>="A=1 B=2"
  envs=$(eval echo \$ENV_$v)
  env $envs SB_LIB_PATH=$PWD/$lib timeout 900 python bench.py --steps ${STEPS:-3} --warmup 3 --no-e2e --cpu-seconds 0.1 ${BENCH_ARGS:---no-sub} \
      >> gpurun_out/abm_$v.json 2>> gpurun_out/abm_$v.err
  rc=$?
  tail -1 gpurun_out/abm_$v.json | python -c "
import json,sys
d=json.loads(sys.stdin.readline())
r=d['roofline']; s=d.get('c2_streaming') or {}; c4=d.get('c4_loop_closure') or {}; c2=d.get('c2_batch') or {}; c3=d.get('c3_knn_normals') or {}
print('$v rc=$rc value %.0f pairs/s  step %.2f ms  stages %s' % (d['value'], d['ms_per_step'], {k: round(v,2) for k,v in r['stages_ms'].items() if v}))
if c2: print('   c2 batch %.2f ms %s | stream mean %.3f p50 %.3f p99 %.3f | c3 %.3f ms | c4 %.3f ms' % (c2.get('ms_per_batch',0), c2.get('stages_ms_last_batch'), s.get('ms_per_frame_mean',0), s.get('p50',0), s.get('p99',0), c3.get('ms',0), c4.get('ms_per_detect',0)))
"
done
done
