#!/bin/bash
# usage: tools/gpurun_retry.sh [gpurun options] -- '<command>' : retries while the pod answers "no slot right now"
for attempt in 1 2 3 4 5 6 7 8 9 10; do
  /usr/local/graft/bin/gpurun "$@" > /tmp/gpurun_retry.out 2>&1
  rc=$?
  if [ $rc -ne 3 ] && ! grep -q "status=transient" /tmp/gpurun_retry.out; then cat /tmp/gpurun_retry.out; exit $rc; fi
  sleep 90
done
cat /tmp/gpurun_retry.out; exit 3
