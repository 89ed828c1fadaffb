#!/bin/bash
# A/B of two builds of the library in one gpurun call: GPU tests on the new build, then bench.py (device-resident C5 step +
# the C2 / C3 / C4 sub-results) on libslam_b200_base.so and on libslam_b200.so, then — new build only — a warm-cache
# launch list of the streaming loop (ncu --cache-control none: kernel durations close to what a frame sees).
mkdir -p gpurun_out
P=lidar-slam-from-scratch_b200
timeout 900 python -m pytest tests -m gpu -q -x --tb=short --timeout 600 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
for v in base new base new; do
  lib=$P/libslam_b200.so; [ $v = base ] && lib=$P/libslam_b200_base.so
  [ -f $lib ] || { echo "no $lib"; continue; }
  SB_LIB_PATH=$PWD/$lib timeout 900 python bench.py --steps ${STEPS:-3} --warmup 3 --no-e2e --cpu-seconds 0.1 ${BENCH_ARGS} \
      >> gpurun_out/ab_$v.json 2>> gpurun_out/ab_$v.err
  echo "$v exit $?"
  tail -1 gpurun_out/ab_$v.json | python -c "
import json,sys
d=json.loads(sys.stdin.readline())
r=d['roofline']; s=d.get('c2_streaming') or {}; c4=d.get('c4_loop_closure') or {}; c2=d.get('c2_batch') or {}; c3=d.get('c3_knn_normals') or {}
print('  value %.0f pairs/s  step %.2f ms  stages %s' % (d['value'], d['ms_per_step'], {k: round(v,2) for k,v in r['stages_ms'].items() if v}))
print('  c2 batch %.2f ms %s | stream mean %.3f p50 %.3f p99 %.3f split %s | c3 %.3f ms | c4 %.3f ms' % (c2.get('ms_per_batch',0), c2.get('stages_ms_last_batch'), s.get('ms_per_frame_mean',0), s.get('p50',0), s.get('p99',0), {k: round(v,3) for k,v in (s.get('split_ms_mean') or {}).items() if v}, c3.get('ms',0), c4.get('ms_per_detect',0)))
"
done
if [ "${NO_NCU:-0}" = "1" ]; then exit 0; fi
export SB_ICP_NOGRAPH=1
CMD="python scripts/stream_latency.py --frames 14"
$CMD > gpurun_out/stream_nograph.json 2> gpurun_out/stream_nograph.err || { echo "stream run failed"; tail -5 gpurun_out/stream_nograph.err; exit 1; }
cat gpurun_out/stream_nograph.json
ncu --metrics gpu__time_duration.sum --cache-control none --clock-control none -c 12000 --csv \
    --log-file gpurun_out/stream_launches_warm.csv $CMD > gpurun_out/ncu_stream.log 2>&1
echo "ncu stream exit $?"
