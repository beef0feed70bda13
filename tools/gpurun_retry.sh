#!/bin/bash
# gpurun_retry.sh TIMEOUT 'command': retries while the pod answers "busy" (exit code 3), every 90 s, up to 40 times
t=$1; shift
for i in $(seq 1 40); do
    /usr/local/graft/bin/gpurun --timeout "$t" -- "$@"
    rc=$?
    [ $rc -ne 3 ] && exit $rc
    sleep 90
done
exit 3
