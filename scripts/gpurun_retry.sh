#!/bin/bash
# Retry a gpurun call while the pool answers "busy" (exit code 3 / status=transient); nothing is charged for those answers.
# usage: scripts/gpurun_retry.sh <log> <gpurun args...>
log=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@" > "$log" 2>&1
  rc=$?
  if ! grep -q "status=transient" "$log"; then exit $rc; fi
  sleep 120
done
exit 3
