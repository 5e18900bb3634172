#!/bin/bash
TAG=$1
tools/_build/microbench_window > gpurun_out/${TAG}_window.json 2>&1
for cs in 12 -1 8 16; do
  FEDDB200_CLASS_CHUNK_SHIFT=$cs timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-ns --no-parity --cpu-M 4 > /tmp/o.json 2>/tmp/o.err
  python -c "
import json; d=json.loads(open('/tmp/o.json').read().strip().splitlines()[-1]); print('chunk_shift $cs', d['ms_per_step'], d['roofline']['frac'])" >> gpurun_out/${TAG}_chunk.log 2>&1
done
cat gpurun_out/${TAG}_chunk.log; cat gpurun_out/${TAG}_window.json
