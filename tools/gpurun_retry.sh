#!/bin/bash
# Retry a gpurun call while the pod answers "busy" (exit code 3). Usage: tools/gpurun_retry.sh <timeout_s> <gpus> '<command>'
t=$1; g=$2; shift 2
for i in $(seq 1 40); do
  if [ "$g" = "1" ]; then /usr/local/graft/bin/gpurun --timeout "$t" -- "$@"; else /usr/local/graft/bin/gpurun --gpus "$g" --timeout "$t" -- "$@"; fi
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 120
done
exit 3
