set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r1e_tests.log
python bench.py --steps 10 --warmup 3 --no-e2e --cpu-M 4 > gpurun_out/r1e_bench_default.json 2> gpurun_out/r1e_bench_default.err
FEDDB200_NO_CLASS_SORT=1 python bench.py --steps 10 --warmup 3 --no-e2e --cpu-M 4 > gpurun_out/r1e_bench_nosort.json 2>&1
FEDDB200_LIB=$PWD/variants/lib_mb2.so python bench.py --steps 10 --warmup 3 --no-e2e --cpu-M 4 > gpurun_out/r1e_bench_mb2.json 2>&1
FEDDB200_GATHER_NT=64 python bench.py --steps 10 --warmup 3 --no-e2e --cpu-M 4 > gpurun_out/r1e_bench_nt64.json 2>&1
cat gpurun_out/r1e_tests.log
for f in default nosort mb2 nt64; do python - $f <<'PY'
import json,sys
f=sys.argv[1]
try:
    d=json.loads(open(f'gpurun_out/r1e_bench_{f}.json').read().strip().splitlines()[-1])
    print(f, round(d["ms_per_step"],4), round(d["roofline"]["frac"],4), d["config"]["pattern_build_s"])
except Exception as e:
    print(f, "FAILED", e)
PY
done
