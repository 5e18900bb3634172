python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r1g_tests.log
cat gpurun_out/r1g_tests.log
python bench.py --steps 10 --warmup 3 --no-e2e --cpu-M 4 > gpurun_out/r1g_bench_default.json 2> gpurun_out/r1g_bench_default.err
FEDDB200_NO_RING=1 python bench.py --steps 10 --warmup 3 --no-e2e --cpu-M 4 > gpurun_out/r1g_bench_noring.json 2>&1
FEDDB200_RING_NT=32 python bench.py --steps 10 --warmup 3 --no-e2e --cpu-M 4 > gpurun_out/r1g_bench_nt32.json 2>&1
FEDDB200_NO_CLASS_SORT=1 python bench.py --steps 10 --warmup 3 --no-e2e --cpu-M 4 > gpurun_out/r1g_bench_nosort.json 2>&1
for f in default noring nt32 nosort; do python - $f <<'PY'
import json,sys
f=sys.argv[1]
try:
    d=json.loads(open(f'gpurun_out/r1g_bench_{f}.json').read().strip().splitlines()[-1])
    print(f, round(d["ms_per_step"],4), round(d["roofline"]["frac"],4), d["config"]["pattern_build_s"])
except Exception as e:
    print(f, "FAILED", e, open(f'gpurun_out/r1g_bench_{f}.json').read()[-500:])
PY
done
bash tools/prof.sh r1g
