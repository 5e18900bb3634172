python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r1f_tests.log
cat gpurun_out/r1f_tests.log
python bench.py --steps 10 --warmup 3 --no-e2e --cpu-M 4 > gpurun_out/r1f_bench_default.json 2> gpurun_out/r1f_bench_default.err
FEDDB200_NO_CLASS_SORT=1 python bench.py --steps 10 --warmup 3 --no-e2e --cpu-M 4 > gpurun_out/r1f_bench_nosort.json 2>&1
FEDDB200_LIB=$PWD/variants/lib_mb4.so python bench.py --steps 10 --warmup 3 --no-e2e --cpu-M 4 > gpurun_out/r1f_bench_mb4.json 2>&1
FEDDB200_LIB=$PWD/variants/lib_mb6.so python bench.py --steps 10 --warmup 3 --no-e2e --cpu-M 4 > gpurun_out/r1f_bench_mb6.json 2>&1
for f in default nosort mb4 mb6; do python - $f <<'PY'
import json,sys
f=sys.argv[1]
try:
    d=json.loads(open(f'gpurun_out/r1f_bench_{f}.json').read().strip().splitlines()[-1])
    print(f, round(d["ms_per_step"],4), round(d["roofline"]["frac"],4), d["config"]["pattern_build_s"])
except Exception as e:
    print(f, "FAILED", e, open(f'gpurun_out/r1f_bench_{f}.json').read()[-500:])
PY
done
bash tools/prof.sh r1f
