#!/bin/bash
# usage: tools/gpurun_retry.sh <log> <timeout_s> [--gpus N] -- '<command>' ; retries while the pod answers "busy" (rc 3, nothing charged)
LOG=$1; shift; TMO=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $TMO "$@" > $LOG 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then echo "gpurun rc=$rc after $i tries"; exit $rc; fi
  sleep 90
done
echo "gave up"; exit 3
