#!/bin/bash
TAG=$1
for fr in 0 1; do
  echo "== FRAG=$fr"
  FEDDB200_FRAG=$fr timeout 600 python tools/bench_ops.py 50 gather 2>&1 | grep -E "laplace |linelas|ns_jac"
  FEDDB200_FRAG=$fr timeout 600 python tools/bench_ops.py 100 gather 2>&1 | grep -E "laplace "
done > gpurun_out/${TAG}_frag_lap.log 2>&1
cat gpurun_out/${TAG}_frag_lap.log
