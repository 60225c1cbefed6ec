#!/bin/bash
# usage: gpurun_retry.sh <timeout_s> <logfile> <command...>   -- retries while the pod answers "busy" (nothing charged)
T=$1; LOG=$2; shift 2
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$T" -- "$@" > "$LOG" 2>&1
  rc=$?
  if grep -q "status=transient" "$LOG" || [ $rc -eq 3 ]; then echo "try $i: busy" >> "$LOG.tries"; sleep 120; continue; fi
  echo "try $i: rc=$rc" >> "$LOG.tries"
  exit $rc
done
exit 3
