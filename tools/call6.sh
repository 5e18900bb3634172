#!/bin/bash
# launch lists (device time per launch) of single operator calls: tools/call6.sh <tag> op...
TAG=$1; shift
for OP in "$@"; do
  python tools/prof_ops.py $OP 50 2 > /dev/null 2>&1
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,launch__registers_per_thread,sm__warps_active.avg.per_cycle_active,smsp__issue_active.avg.per_cycle_active,smsp__inst_executed.sum,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active --clock-control none --csv --log-file gpurun_out/${TAG}_launch_$OP.csv python tools/prof_ops.py $OP 50 2 > gpurun_out/${TAG}_launch_$OP.log 2>&1
done
