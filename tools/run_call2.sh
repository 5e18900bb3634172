#!/bin/bash
# usage: tools/run_call2.sh <tag> -- GPU tests, then bench lines (elasticity only) under the env settings listed in $ENVS (';'-separated)
TAG=$1
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 ) > gpurun_out/${TAG}_tests.log
tail -5 gpurun_out/${TAG}_tests.log
IFS=';' read -ra E <<< "$ENVS"
i=0
for e in "${E[@]}"; do
  i=$((i+1))
  ( env $e timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-ns --cpu-M 4 > gpurun_out/${TAG}_b$i.json 2> gpurun_out/${TAG}_b$i.err )
  python - "$e" gpurun_out/${TAG}_b$i.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    print(sys.argv[1], "ms", round(d["ms_per_step"],4), "frac", round(d["roofline"]["frac"],4), "chk", d["checksum_first_1Mi_values"])
except Exception as ex:
    print(sys.argv[1], "FAILED", ex, open(sys.argv[2].replace(".json",".err")).read()[-600:])
PY
done
