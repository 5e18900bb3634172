#!/bin/bash
# usage: tools/ab2.sh tag "ENV=.. ENV=.." ...   -- one bench line per configuration (tuning aid)
TAG=$1; shift
i=0
for cfg in "$@"; do
  i=$((i+1))
  env $cfg python bench.py --steps 10 --warmup 3 --no-e2e --no-ns --cpu-M 4 > gpurun_out/${TAG}_$i.json 2> gpurun_out/${TAG}_$i.err
  python - "$cfg" gpurun_out/${TAG}_$i.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    print(sys.argv[1], '|', round(d["ms_per_step"],4), round(d["roofline"]["frac"],4), round(d["config"]["pattern_build_s"],2))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
