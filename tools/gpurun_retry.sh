#!/bin/bash
# gpurun with retries while the pod answers "busy" (exit 3: nothing charged). Usage: tools/gpurun_retry.sh [gpurun args] -- 'cmd'
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun "$@"
  rc=$?
  [ $rc -ne 3 ] && exit $rc
  sleep 45
done
exit 3
