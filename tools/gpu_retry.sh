#!/bin/bash
# usage: tools/gpu_retry.sh <timeout> <command...>   — retries while the pod answers "busy" (exit code 3)
t=$1; shift
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun --timeout $t "$@"; rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 60
done
exit 3
