#!/bin/bash
# usage: tools/gpurun_retry.sh [gpurun args...]  -- retries while gpurun answers "busy" (exit 3)
for attempt in $(seq 1 30); do
    /usr/local/graft/bin/gpurun "$@"
    rc=$?
    if [ $rc -ne 3 ]; then exit $rc; fi
    echo "[retry] busy (attempt $attempt), sleeping 120 s" >&2
    sleep 120
done
exit 3
