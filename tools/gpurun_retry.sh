#!/bin/bash
# Developer helper: retry a gpurun call while the pod answers "busy" (exit code 3, nothing charged).
# usage: tools/gpurun_retry.sh <timeout-seconds> '<command>' [extra gpurun flags]
T=$1; CMD=$2; shift 2
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$T" "$@" -- "$CMD"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 45
done
exit 3
