#!/bin/bash
# Produces the ncu evidence of one round (run under gpurun; copy the summaries from gpurun_out/ into profiles/).
#   1. plain run of the bench command (must exit 0), 2. launch list with device times, 3. --set full capture of one step
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-ns --cpu-M 4"
$CMD > gpurun_out/prof_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 0 -c 400 --csv --log-file gpurun_out/prof_launches.csv $CMD > gpurun_out/prof_launches.log 2>&1
$CMD > gpurun_out/prof_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"k_gather|k_ring|k_geom" -s 33 -c 11 -f -o gpurun_out/prof_step $CMD > gpurun_out/prof_step.log 2>&1
ncu -i gpurun_out/prof_step.ncu-rep --page raw --csv > gpurun_out/prof_step_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_step.ncu-rep --page source --csv > gpurun_out/prof_step_src.csv 2>/dev/null
tail -2 gpurun_out/prof_step.log; tail -1 gpurun_out/prof_plain.log | cut -c1-300
