#!/bin/bash
# Retry `gpurun` while the pool answers "busy" (exit code 3: nothing charged).  Usage: tools/gpurun_retry.sh [gpurun args] -- '<command>'
for attempt in $(seq 1 12); do
    /usr/local/graft/bin/gpurun "$@"
    rc=$?
    if [ $rc -ne 3 ]; then exit $rc; fi
    echo "[retry] attempt $attempt answered busy; sleeping 60 s" >&2
    sleep 60
done
exit 3
