#!/bin/bash
# Produces the ncu evidence of one round (run under gpurun; summaries are copied from gpurun_out/ into profiles/ afterwards).
#   1. plain run of the bench command (must exit 0), 2. launch list with device times, 3. --set full capture of one step
TAG=${1:-r02}
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-ns --no-parity --cpu-M 4"
$CMD > gpurun_out/${TAG}_prof_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 0 -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_launches.log 2>&1
$CMD > gpurun_out/${TAG}_prof_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"k_gather|k_ring|k_geom|k_task" -s 30 -c 10 -f -o gpurun_out/${TAG}_step $CMD > gpurun_out/${TAG}_step.log 2>&1
ncu -i gpurun_out/${TAG}_step.ncu-rep --page raw --csv > gpurun_out/${TAG}_step_raw.csv 2>/dev/null
ncu -i gpurun_out/${TAG}_step.ncu-rep --page source --csv > gpurun_out/${TAG}_step_src.csv 2>/dev/null
tail -2 gpurun_out/${TAG}_step.log; tail -1 gpurun_out/${TAG}_prof_plain.log | cut -c1-300
# 4. Navier-Stokes block (fused (0,0) block on the structured P2 cube M=50): launch list + --set full capture of its kernels
OPC="python tools/prof_ops.py nsj 50 2"
$OPC > gpurun_out/${TAG}_ns_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"k_gatherw|k_sloc|k_udata|k_gatherx" -s 13 -c 6 -f -o gpurun_out/${TAG}_ns $OPC > gpurun_out/${TAG}_ns.log 2>&1
ncu -i gpurun_out/${TAG}_ns.ncu-rep --page raw --csv > gpurun_out/${TAG}_ns_raw.csv 2>/dev/null
tail -2 gpurun_out/${TAG}_ns.log
