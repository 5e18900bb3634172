#!/bin/bash
TAG=$1; N=$2
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29502 bench.py --gpus $N --steps 10 --warmup 3 --no-ns --no-e2e --cpu-M 4 > gpurun_out/${TAG}_b${N}.json 2> gpurun_out/${TAG}_b${N}.err; echo "bench rc=$?" >> gpurun_out/${TAG}_b${N}.err
tail -3 gpurun_out/${TAG}_b${N}.err
python - <<PY
import json
d=json.loads(open("gpurun_out/${TAG}_b${N}.json").read().strip().splitlines()[-1])
print(d["n_gpus"], "ms", round(d["ms_per_step"],4), "ns_mgpu", d.get("navier_stokes_block_multi_gpu"), "parity", (d.get("parity_check") or {}).get("all_ranks"))
PY
