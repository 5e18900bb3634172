#!/bin/bash
# A/B of the concurrent ghost rows: tools/call_mgpu2.sh <tag> <N>
TAG=$1; N=$2
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
T_M=4 timeout 600 $TR --master-port 29501 tests/dist_gpu_check.py > gpurun_out/${TAG}_dist${N}.log 2>&1; echo "dist_check rc=$?" >> gpurun_out/${TAG}_dist${N}.log
for cg in 1 0 1 0; do
  FEDDB200_CONCURRENT_GHOST=$cg timeout 900 $TR --master-port 2951$cg bench.py --gpus $N --steps 20 --warmup 3 --no-ns --no-e2e --cpu-M 4 > /tmp/b.json 2> /tmp/b.err
  python - <<PY >> gpurun_out/${TAG}_cg${N}.log
import json
try:
    d=json.loads(open("/tmp/b.json").read().strip().splitlines()[-1])
    print("concurrent_ghost=$cg", d["n_gpus"], "ms", round(d["ms_per_step"],4), "parity", (d.get("parity_check") or {}).get("all_ranks"), d.get("multi_gpu_phase_ms"))
except Exception as e: print("ERR", e, open("/tmp/b.err").read()[-600:])
PY
done
tail -3 gpurun_out/${TAG}_dist${N}.log; cat gpurun_out/${TAG}_cg${N}.log
