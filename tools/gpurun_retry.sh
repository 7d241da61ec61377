#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout_s> <script> <logfile>   — retries while the pod answers "busy / draining" (rc 3)
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$1" -- "bash $2" > "$3" 2>&1
  rc=$?
  if grep -q "status=transient\|status=busy" "$3" || [ $rc -eq 3 ]; then sleep 150; continue; fi
  break
done
