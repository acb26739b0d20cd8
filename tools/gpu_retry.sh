#!/bin/bash
# usage: [GPUS=8] tools/gpu_retry.sh <timeout-seconds> '<command>'   -- retries gpurun while the pod answers busy (exit 3 / transient)
T=$1; shift
for i in $(seq 1 40); do
  out=$(/usr/local/graft/bin/gpurun ${GPUS:+--gpus $GPUS} --timeout "$T" -- "$@" 2>&1)
  rc=$?
  if echo "$out" | grep -q "status=transient"; then sleep 90; continue; fi
  if [ $rc -eq 3 ]; then sleep 90; continue; fi
  echo "$out"
  exit $rc
done
echo "gave up after 40 tries"
exit 3
