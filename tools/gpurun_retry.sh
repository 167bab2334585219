#!/bin/bash
# gpurun with retries while the pod answers busy (exit 3): tools/gpurun_retry.sh <log> <timeout> <command...>
log=$1; shift; to=$1; shift
for i in 1 2 3 4 5 6 7 8; do
  gpurun --timeout "$to" -- "$@" > "$log" 2>&1
  rc=$?
  if ! grep -q "status=transient" "$log"; then exit $rc; fi
  sleep 120
done
exit 3
