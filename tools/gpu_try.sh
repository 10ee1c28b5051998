#!/bin/bash
# usage: tools/gpu_try.sh <timeout> <gpus> '<command>'   -- retries gpurun while the pod answers busy (exit 3 / transient)
T=$1; G=$2; shift 2
for i in $(seq 1 30); do
  if [ "$G" = "1" ]; then out=$(/usr/local/graft/bin/gpurun --timeout $T -- "$@" 2>&1); else out=$(/usr/local/graft/bin/gpurun --gpus $G --timeout $T -- "$@" 2>&1); fi
  rc=$?
  if echo "$out" | grep -q "nothing was charged"; then sleep 45; continue; fi
  echo "$out"; exit $rc
done
echo "gave up: pod busy"; exit 3
