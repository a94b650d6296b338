#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout> <script> [gpus]  - retries while the pod answers "transient" (nothing charged)
T=$1; S=$2; G=${3:-1}
for i in $(seq 1 40); do
  if [ "$G" = "1" ]; then out=$(/usr/local/graft/bin/gpurun --timeout $T -- "bash $S" 2>&1); else out=$(/usr/local/graft/bin/gpurun --gpus $G --timeout $T -- "bash $S" 2>&1); fi
  if echo "$out" | grep -q "status=transient"; then sleep 90; continue; fi
  echo "$out" | tail -60; exit 0
done
echo "gave up after 40 transient answers"
