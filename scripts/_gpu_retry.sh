#!/bin/bash
# usage: _gpu_retry.sh <timeout> <script>   -- resubmit while the pod answers "busy" (nothing is charged for those)
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun --timeout "$1" -- "bash $2" > /tmp/gpu_retry.out 2>&1
  if grep -q "status=transient\|exit code 3\|rc=3" /tmp/gpu_retry.out && ! grep -q "status=ok" /tmp/gpu_retry.out; then sleep 100; continue; fi
  break
done
tail -${3:-40} /tmp/gpu_retry.out
