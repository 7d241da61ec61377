#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout_s> <script> <logfile> [gpus]   — retries while the pod answers "busy / draining" (nothing is charged then)
GP=""; if [ -n "$4" ]; then GP="--gpus $4"; fi
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun $GP --timeout "$1" -- "bash $2" > "$3" 2>&1
  rc=$?
  if grep -q "status=transient\|status=busy" "$3" || [ $rc -eq 3 ]; then sleep 120; continue; fi
  break
done
