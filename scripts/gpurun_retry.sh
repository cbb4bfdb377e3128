#!/bin/bash
# gpurun with retries while the pod answers "busy" (exit code 3, nothing charged): gpurun_retry.sh LOG TIMEOUT 'command'
log=$1; to=$2; shift 2
for i in $(seq 1 ${RETRIES:-15}); do
  /usr/local/graft/bin/gpurun ${GPUS:+--gpus $GPUS} --timeout $to -- "$@" > $log 2>&1
  rc=$?
  [ $rc -ne 3 ] && exit $rc
  sleep ${RETRY_SLEEP:-90}
done
exit 3
