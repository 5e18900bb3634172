#!/bin/bash
# elasticity (M=70) and scalar Laplace (M=100) for variant libraries, one box
TAG=$1; shift
for V in default "$@"; do
  if [ "$V" != "default" ]; then export FEDDB200_LIB=$PWD/variants/lib_$V.so; else unset FEDDB200_LIB; fi
  echo "== $V"
  timeout 300 python tools/bench_ops.py 70 gather 2>&1 | grep -E "linelas"
  timeout 300 python tools/bench_ops.py 100 gather 2>&1 | grep -E "laplace "
done > gpurun_out/${TAG}_ab3.log 2>&1
cat gpurun_out/${TAG}_ab3.log
