#!/bin/bash
# usage: gpurun_retry.sh <timeout> [--gpus N] -- '<command>' : retries while the pod answers "busy" (exit 3)
T=$1; shift
for k in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$T" "$@" > /tmp/gpurun_last.log 2>&1
  rc=$?
  if grep -q "status=transient" /tmp/gpurun_last.log || [ $rc -eq 3 ]; then sleep 60; continue; fi
  break
done
cat /tmp/gpurun_last.log
exit $rc
