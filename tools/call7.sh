#!/bin/bash
TAG=$1; shift
( timeout 900 python -m pytest tests -m gpu -x -q -k "advection or ns_jac or interface or cpp" 2>&1 | tail -3 ) > gpurun_out/${TAG}_tests.log
cat gpurun_out/${TAG}_tests.log
for V in "" "$@"; do
  if [ -n "$V" ]; then export FEDDB200_LIB=$PWD/variants/lib_$V.so; fi
  timeout 600 python tools/bench_ops.py 50 > gpurun_out/${TAG}_ops_${V:-default}.log 2>&1
  echo "== ${V:-default}"; cat gpurun_out/${TAG}_ops_${V:-default}.log
done
