#!/bin/bash
# usage: tools/gpurun_retry.sh LOGFILE [gpurun args...] ; retries while the pod answers busy/transient (exit 3)
LOG=$1; shift
for i in $(seq 1 12); do
  /usr/local/graft/bin/gpurun "$@" > "$LOG" 2>&1
  rc=$?
  if grep -q "status=transient\|retry in a few minutes" "$LOG" || [ $rc -eq 3 ]; then sleep 150; continue; fi
  exit $rc
done
exit 3
