#!/bin/bash
# usage: tools/gpurun_retry.sh LOGFILE [gpurun args...] -- command     (retries while the pod answers "transient"/busy)
log="$1"; shift
for attempt in 1 2 3 4 5 6 7 8 9 10; do
  /usr/local/graft/bin/gpurun "$@" > "$log" 2>&1
  rc=$?
  if grep -q "status=transient" "$log" || [ $rc -eq 3 ]; then sleep 150; continue; fi
  break
done
