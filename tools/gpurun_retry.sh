#!/bin/bash
# gpurun with retries while the pod answers "busy" (exit code 3: nothing charged).  usage: tools/gpurun_retry.sh [gpurun args] -- 'command'
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@" > gpurun_out/.retry.log 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then cat gpurun_out/.retry.log; exit $rc; fi
  sleep 90
done
cat gpurun_out/.retry.log; exit 3
