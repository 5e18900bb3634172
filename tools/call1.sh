#!/bin/bash
# one gpurun call: quick parity subset, A/B lines of the variant libraries given as arguments, the default line
TAG=$1; shift
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "elas or linelas or laplace" 2>&1 | tail -5 ) > gpurun_out/${TAG}_tests.log
for V in "$@"; do tools/ab.sh $V > gpurun_out/${TAG}_ab_$V.log 2>&1; done
timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-ns --cpu-M 4 > gpurun_out/${TAG}_default.json 2> gpurun_out/${TAG}_default.err
tail -3 gpurun_out/${TAG}_tests.log; cat gpurun_out/${TAG}_ab_*.log; python -c "
import json; d=json.loads(open('gpurun_out/${TAG}_default.json').read().strip().splitlines()[-1]); print('default', d['ms_per_step'], d['roofline']['frac'], d.get('parity_check'))"; tail -3 gpurun_out/${TAG}_default.err
