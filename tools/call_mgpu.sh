#!/bin/bash
# multi-GPU call: tools/call_mgpu.sh <tag> <N> [extra bench args]   (run under gpurun --gpus N)
TAG=$1; N=$2; shift 2
( nvidia-smi topo -m; lscpu | grep -i -E "numa|socket|model name|^CPU\(s\)"; for d in /sys/bus/pci/devices/*; do if [ -f $d/numa_node ] && grep -q 0x10de $d/vendor 2>/dev/null; then echo "$d $(cat $d/numa_node) $(cat $d/class)"; fi; done; cat /sys/devices/system/node/online 2>/dev/null; taskset -p $$; free -g | head -2 ) > gpurun_out/${TAG}_topo${N}.log 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
T_M=4 timeout 600 $TR --master-port 29501 tests/dist_gpu_check.py > gpurun_out/${TAG}_dist${N}.log 2>&1; echo "dist_check rc=$?" >> gpurun_out/${TAG}_dist${N}.log
timeout 900 $TR --master-port 29502 bench.py --gpus $N --steps 10 --warmup 3 --no-ns "$@" > gpurun_out/${TAG}_b${N}.json 2> gpurun_out/${TAG}_b${N}.err; echo "bench rc=$?" >> gpurun_out/${TAG}_b${N}.err
timeout 900 $TR --master-port 29503 bench.py --gpus $N --steps 5 --warmup 3 --no-ns --no-e2e --M 101 --cpu-M 4 > gpurun_out/${TAG}_b${N}_m101.json 2> gpurun_out/${TAG}_b${N}_m101.err; echo "bench101 rc=$?" >> gpurun_out/${TAG}_b${N}_m101.err
tail -4 gpurun_out/${TAG}_dist${N}.log; tail -3 gpurun_out/${TAG}_b${N}.err; tail -2 gpurun_out/${TAG}_b${N}_m101.err
python - <<PY
import json
for f in ("gpurun_out/${TAG}_b${N}.json","gpurun_out/${TAG}_b${N}_m101.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["n_gpus"], "ms", round(d["ms_per_step"],4), "frac", round(d["roofline"]["frac"],4), "e2e", d.get("e2e"), "phases", d.get("multi_gpu_phase_ms"), "parity", d.get("parity_check"), "build", d["config"].get("pattern_build_s"), d["config"].get("pattern_build_breakdown"))
    except Exception as e: print(f, "ERR", e)
PY
