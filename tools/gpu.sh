#!/bin/bash
# usage: tools/gpu.sh <timeout-seconds> <gpus> <command...>   -- retries while the pod answers "busy" (nothing charged)
T=$1; G=$2; shift 2
for i in $(seq 1 40); do
  out=$(/usr/local/graft/bin/gpurun --gpus $G --timeout $T -- "$@" 2>&1)
  if echo "$out" | grep -q "status=transient"; then sleep 45; continue; fi
  echo "$out"; exit 0
done
echo "gpu.sh: still busy after 40 tries"; exit 3
