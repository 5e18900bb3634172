#!/bin/bash
# usage: tools/prof.sh <tag> <kernel-regex> <skip> <count> [extra bench args]  -- plain run, then ncu --set full of some row-kernel launches
TAG=$1; RE=$2; SKIP=$3; CNT=$4; shift 4
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-ns --cpu-M 4 $@"
$CMD > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"$RE" -s $SKIP -c $CNT -f -o gpurun_out/prof_$TAG $CMD > gpurun_out/ncu_$TAG.log 2>&1
ncu -i gpurun_out/prof_$TAG.ncu-rep --page raw --csv > gpurun_out/raw_$TAG.csv 2>/dev/null
ncu -i gpurun_out/prof_$TAG.ncu-rep --page source --csv > gpurun_out/src_$TAG.csv 2>/dev/null
tail -2 gpurun_out/ncu_$TAG.log
