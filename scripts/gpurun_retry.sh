#!/bin/bash
# usage: scripts/gpurun_retry.sh <timeout_s> <job script> [--gpus N]   retries while the pod answers "transient"/busy
T=$1; JOB=$2; shift 2
for i in $(seq 1 30); do
  OUT=$(/usr/local/graft/bin/gpurun --timeout $T "$@" -- "bash $JOB" 2>&1)
  echo "$OUT" | tail -70
  if echo "$OUT" | grep -q "status=transient\|exit code 3\|status=busy"; then sleep 90; continue; fi
  break
done
