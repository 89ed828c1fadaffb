#!/bin/bash
# gpurun with retries while the pod answers "transient"/busy (exit 3 or status=transient): usage [GPUS=N] gpurun_retry.sh <timeout> <cmd>
T=$1; shift
for i in $(seq 1 20); do
  out=$(/usr/local/graft/bin/gpurun --timeout $T ${GPUS:+--gpus $GPUS} -- "$@" 2>&1); rc=$?
  echo "$out" | tail -${TAILN:-25}
  if echo "$out" | grep -q "status=transient\|status=busy" || [ $rc -eq 3 ]; then sleep 60; continue; fi
  exit $rc
done
