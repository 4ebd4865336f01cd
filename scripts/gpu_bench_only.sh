#!/bin/bash
# default bench line only. Usage: gpurun -- 'bash scripts/gpu_bench_only.sh TAG [bench args]'
TAG=${1:-b}; shift
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python bench.py "$@" > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench exit $?"
tail -3 $OUT/bench_$TAG.err
python - <<PY
import json
d=json.loads(open("$OUT/bench_$TAG.json").read().strip().splitlines()[-1])
print("ms/frame", d["ms_per_step"], "Mpix/s", d["value"], "e2e", d["e2e"])
print(d.get("pass_ms"))
PY
