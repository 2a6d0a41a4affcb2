#!/bin/bash
# Retry a gpurun call while the pod answers "no slot right now" (exit code 3; nothing is charged).
# usage: tools/gpurun_retry.sh [gpurun options] -- 'command'
for attempt in $(seq 1 30); do
    /usr/local/graft/bin/gpurun "$@"
    rc=$?
    if [ $rc -ne 3 ]; then exit $rc; fi
    echo "[gpurun_retry] attempt $attempt: no slot, retrying in 60 s" >&2
    sleep 60
done
exit 3
