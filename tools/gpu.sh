#!/bin/bash
# gpurun with retries while the pod has no free slot (exit 3 / transient).
# usage: tools/gpu.sh <timeout_s> <logname> '<command>'
T=$1; NAME=$2; shift 2
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$T" -- "$@" > gpurun_out/$NAME.txt 2>&1
  rc=$?
  if grep -q "status=transient" gpurun_out/$NAME.txt || [ $rc -eq 3 ]; then sleep 60; continue; fi
  exit $rc
done
exit 3
