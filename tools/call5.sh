#!/bin/bash
TAG=$1
( timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 ) > gpurun_out/${TAG}_tests.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-e2e --cpu-M 4 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
cat gpurun_out/${TAG}_tests.log; tail -3 gpurun_out/${TAG}_bench.err
python - <<PY
import json
d=json.loads(open("gpurun_out/${TAG}_bench.json").read().strip().splitlines()[-1])
print("bench", d["ms_per_step"], d["roofline"]["frac"])
for k in ("navier_stokes_blocks","config4_navier_stokes_dfg3d","laplace_configs_1_2"):
    v=d.get(k) or {}
    print(k, {kk:(round(vv["ms"],3), round(vv.get("hbm_frac",0),3)) for kk,vv in v.items() if isinstance(vv,dict) and "ms" in vv})
PY
