#!/bin/bash
# Retry a gpurun call while the pod answers "busy" (exit 3: nothing charged).  usage: tools/gpu_retry.sh <log> <gpurun args...>
LOG=$1; shift
for i in $(seq 1 30); do
    /usr/local/graft/bin/gpurun "$@" > "$LOG" 2>&1
    rc=$?
    if ! grep -q "status=transient" "$LOG"; then exit $rc; fi
    sleep 90
done
exit 3
