#!/bin/bash
# usage: scripts/gpurun_retry.sh [--gpus N] <timeout> <script>   retries while the pod answers "transient" (nothing charged)
gp=""; if [ "$1" = "--gpus" ]; then gp="--gpus $2"; shift 2; fi
for i in 1 2 3 4 5 6 7 8 9 10 11 12; do
  /usr/local/graft/bin/gpurun $gp --timeout "$1" -- "$2" > /tmp/gpurun_retry.out 2>&1
  if grep -q "status=transient" /tmp/gpurun_retry.out; then sleep 100; continue; fi
  break
done
cat /tmp/gpurun_retry.out
