#!/bin/bash
# one gpurun call: GPU tests, pipe-rate micro-benchmarks, the default bench line and A/B lines of variant libraries
# usage: tools/run_call.sh <tag> [variant ...]
TAG=$1; shift
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 ) > gpurun_out/${TAG}_tests.log
tools/_build/microbench > gpurun_out/${TAG}_microbench.json 2> gpurun_out/${TAG}_microbench.err
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
for V in "$@"; do
  tools/ab.sh $V > gpurun_out/${TAG}_ab_$V.log 2>&1
done
tail -3 gpurun_out/${TAG}_tests.log; cat gpurun_out/${TAG}_microbench.json; cat gpurun_out/${TAG}_ab_*.log
python - <<PY
import json
d=json.loads(open("gpurun_out/${TAG}_bench.json").read().strip().splitlines()[-1])
print("bench", d["ms_per_step"], d["roofline"]["frac"], d["e2e"], {k:(v["ms"] if isinstance(v,dict) else v) for k,v in (d.get("navier_stokes_blocks") or {}).items()})
PY
