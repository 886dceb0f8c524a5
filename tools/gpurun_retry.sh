#!/bin/bash
# tools/gpurun_retry.sh <timeout> <logfile> <command...>: gpurun, retried while the pod answers "busy" (exit 3)
to=$1; log=$2; shift 2
for i in $(seq 1 40); do
    /usr/local/graft/bin/gpurun --timeout $to -- "$@" > $log 2>&1
    rc=$?
    [ $rc -ne 3 ] && exit $rc
    sleep 150
done
exit 3
