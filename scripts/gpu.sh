#!/bin/bash
# gpurun with retries while the pod has no free slot (nothing is charged for those):  scripts/gpu.sh <timeout> <log> <command...>
T=$1; LOG=$2; shift 2
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun --timeout $T -- "$@" > $LOG 2>&1
  if grep -q "status=transient\|status=busy\|exit code 3" $LOG; then sleep 90; else break; fi
done
grep -v "merged" $LOG | tail -40
