#!/bin/bash
# tools/gpurun_slim.sh in a retry loop while the pod answers "transient" (no slot free; nothing charged)
# usage: [KEEP=...] tools/gpurun_retry.sh <timeout_s> '<command>' <logfile>
for i in $(seq 1 40); do
  "$(dirname "$0")/gpurun_slim.sh" "$1" "$2" > "$3" 2>&1
  grep -q "status=transient" "$3" || break
  sleep 90
done
