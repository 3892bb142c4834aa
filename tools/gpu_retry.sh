#!/bin/bash
# usage: tools/gpu_retry.sh <timeout_s> <gpus> <script>   — retries while the pod answers "busy" (nothing is charged for those)
T=$1; G=$2; S=$3
for i in $(seq 1 40); do
  if [ "$G" = "1" ]; then OUT=$(/usr/local/graft/bin/gpurun --timeout $T -- "bash $S" 2>&1); else OUT=$(/usr/local/graft/bin/gpurun --gpus $G --timeout $T -- "bash $S" 2>&1); fi
  if echo "$OUT" | grep -q "status=transient"; then sleep 60; continue; fi
  echo "$OUT"; exit 0
done
echo "gave up"; exit 3
