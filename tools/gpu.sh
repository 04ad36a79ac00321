#!/bin/bash
# gpurun with retries while the pod has no free slot (exit 3 = nothing charged).  usage: tools/gpu.sh <timeout_s> '<command>'
t=$1; shift
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun --timeout "$t" -- "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 45
done
exit 3
