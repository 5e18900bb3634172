#!/bin/bash
TAG=$1
for cfg in "default 128" "default 64" "su2 128" "su2 64"; do
  set -- $cfg
  if [ "$1" != "default" ]; then export FEDDB200_LIB=$PWD/variants/lib_$1.so; else unset FEDDB200_LIB; fi
  echo "== $cfg"
  FEDDB200_SLOC_NT=$2 timeout 600 python tools/bench_ops.py 50 gather 2>&1 | grep -E "advection N|ns_jac"
done > gpurun_out/${TAG}_sloc.log 2>&1
cat gpurun_out/${TAG}_sloc.log
