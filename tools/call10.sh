#!/bin/bash
TAG=$1
( timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 ) > gpurun_out/${TAG}_tests.log
cat gpurun_out/${TAG}_tests.log
for pt in 1 0; do
  echo "== LAPLACE_POINTS=$pt"
  FEDDB200_LAPLACE_POINTS=$pt timeout 600 python tools/bench_ops.py 50 gather 2>&1 | grep -E "laplace|linelas"
  FEDDB200_LAPLACE_POINTS=$pt timeout 600 python tools/bench_ops.py 100 gather 2>&1 | grep -E "laplace |laplace_vec"
done > gpurun_out/${TAG}_pts.log 2>&1
cat gpurun_out/${TAG}_pts.log
