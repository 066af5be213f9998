#!/bin/bash
# Local helper (build container): call gpurun, retrying while the pod answers "busy" (exit code 3: nothing charged).
# usage: tools/gpurun_retry.sh LOGFILE [gpurun args...]
log=$1; shift
for attempt in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@" > "$log" 2>&1
  rc=$?
  if [ $rc -ne 3 ] && ! grep -q "status=transient" "$log"; then exit $rc; fi
  sleep 120
done
exit 3
