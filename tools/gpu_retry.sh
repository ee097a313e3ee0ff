#!/bin/bash
# gpurun wrapper for the build container: retries while the pod answers "no box free right now" (exit 3).
for i in $(seq 1 60); do
    /usr/local/graft/bin/gpurun "$@"
    rc=$?
    [ $rc -ne 3 ] && exit $rc
    sleep 75
done
exit 3
