#!/bin/bash
# usage: tools/gpu_retry.sh <log> <timeout_s> <command...>   - retries gpurun while the pod answers busy (exit 3)
LOG=$1; TO=$2; shift 2
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun --timeout "$TO" -- "$@" > "$LOG" 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 120
done
exit 3
