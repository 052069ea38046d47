#!/bin/bash
# usage: scripts/gpuretry.sh <gpurun args...>   - retries while the pod answers "busy" (exit 3)
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  echo "[gpuretry] busy, attempt $i; sleeping 90 s"
  sleep 90
done
exit 3
