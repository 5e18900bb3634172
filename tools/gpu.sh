#!/bin/bash
# retry wrapper around gpurun (exit 3 = no slot right now, nothing charged): tools/gpu.sh <log> <timeout-s> [--gpus N] -- '<command>'
LOG=$1; TMO=$2; shift 2
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $TMO "$@" > $LOG 2>&1
  rc=$?
  if grep -q "status=transient" $LOG || [ $rc -eq 3 ]; then sleep 90; continue; fi
  break
done
echo "gpu.sh done rc=$rc" >> $LOG
