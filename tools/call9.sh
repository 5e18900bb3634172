#!/bin/bash
TAG=$1; V=$2
export FEDDB200_LIB=$PWD/variants/lib_$V.so
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-ns --no-parity --cpu-M 4"
$CMD > gpurun_out/${TAG}_plain.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_ring" -s 6 -c 2 -f -o gpurun_out/${TAG}_whatif $CMD > gpurun_out/${TAG}_ncu.log 2>&1
ncu -i gpurun_out/${TAG}_whatif.ncu-rep --page raw --csv > gpurun_out/${TAG}_whatif_raw.csv 2>/dev/null
ncu -i gpurun_out/${TAG}_whatif.ncu-rep --page source --csv > gpurun_out/${TAG}_whatif_src.csv 2>/dev/null
tail -2 gpurun_out/${TAG}_ncu.log | cut -c1-200
