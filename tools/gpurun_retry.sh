#!/bin/bash
# gpurun with retries while the pod has no free GPU slot (exit code 3 = nothing charged).  usage: gpurun_retry.sh <gpurun args...>
for attempt in $(seq 1 20); do
  /usr/local/graft/bin/gpurun "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  echo "[gpurun_retry] busy (attempt $attempt), sleeping 90 s"
  sleep 90
done
exit 3
