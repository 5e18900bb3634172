#!/bin/bash
# A/B of variant libraries on the operator timings: tools/call_ops_ab.sh <tag> <M> <variant> ...
TAG=$1; M=$2; shift 2
for V in default "$@" default "$@"; do
  if [ "$V" != "default" ]; then export FEDDB200_LIB=$PWD/variants/lib_$V.so; else unset FEDDB200_LIB; fi
  echo "== $V"
  timeout 600 python tools/bench_ops.py $M gather 2>&1 | grep -E "advection|ns_jac"
done > gpurun_out/${TAG}_ops_ab.log 2>&1
cat gpurun_out/${TAG}_ops_ab.log
