#!/bin/bash
# usage: gpurun_retry.sh <gpurun args...>   -- retries while the pod answers "busy / transient" (exit code 3), up to ~40 min
for attempt in $(seq 1 14); do
    /usr/local/graft/bin/gpurun "$@" > /tmp/gpurun_last.log 2>&1
    rc=$?
    if grep -q "status=transient" /tmp/gpurun_last.log || [ $rc -eq 3 ]; then
        sleep 150
        continue
    fi
    tail -40 /tmp/gpurun_last.log
    exit $rc
done
tail -5 /tmp/gpurun_last.log
exit 3
