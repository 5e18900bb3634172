#!/bin/bash
# what-if timing variants (wrong values by construction): one line each
TAG=$1; shift
for V in "$@"; do
  FEDDB200_LIB=$PWD/variants/lib_$V.so timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-ns --no-parity --cpu-M 4 > /tmp/o.json 2>/tmp/o.err
  python -c "
import json; d=json.loads(open('/tmp/o.json').read().strip().splitlines()[-1]); print('$V', d['ms_per_step'], d['roofline']['frac'])" >> gpurun_out/${TAG}_whatif.log 2>&1
done
cat gpurun_out/${TAG}_whatif.log
