#!/bin/bash
# usage: tools/ab.sh <variant-lib-suffix> [NT]   -- one bench line for a variant library (tuning aid)
V=$1; NT=$2
if [ -n "$NT" ]; then export FEDDB200_GATHER_NT=$NT; else unset FEDDB200_GATHER_NT; fi
FEDDB200_LIB=$PWD/variants/lib_$V.so timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-ns --cpu-M 4 > /tmp/ab.out 2> /tmp/ab.err
python - "$V" "$NT" <<'PY'
import json,sys
try:
    d=json.loads(open('/tmp/ab.out').read().strip().splitlines()[-1])
    print(sys.argv[1:], round(d["ms_per_step"],4), round(d["roofline"]["frac"],4))
except Exception as e:
    print(sys.argv[1:], "FAILED", open('/tmp/ab.err').read()[-400:])
PY
