#!/bin/bash
# full GPU tests + frag on/off A/B
TAG=$1
( timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 ) > gpurun_out/${TAG}_tests.log
for fr in 1 0; do
  FEDDB200_FRAG=$fr timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-ns --cpu-M 4 > /tmp/o.json 2>/tmp/o.err
  python -c "
import json; d=json.loads(open('/tmp/o.json').read().strip().splitlines()[-1]); print('frag $fr', d['ms_per_step'], d['roofline']['frac'], d['gpu_launches'], (d.get('parity_check') or {}).get('result'), (d.get('parity_check') or {}).get('rel_frobenius'))" >> gpurun_out/${TAG}_frag.log 2>&1
  tail -2 /tmp/o.err >> gpurun_out/${TAG}_frag.log
done
cat gpurun_out/${TAG}_tests.log gpurun_out/${TAG}_frag.log
