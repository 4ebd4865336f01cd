#!/bin/bash
# quick GPU iteration: parity tests + 4K bench line(s). Usage: gpurun -- 'bash scripts/gpu_quick.sh TAG [pytest-args]'
TAG=${1:-q}; shift
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q "$@" > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu_$TAG.log
tail -15 $OUT/pytest_gpu_$TAG.log
timeout 600 python bench.py --steps 100 --no-cpu-baseline > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench exit $?"
python - <<PY
import json
d=json.loads(open("$OUT/bench_$TAG.json").read().strip().splitlines()[-1])
print("ms/frame", d["ms_per_step"], "Mpix/s", d["value"], "e2e", d["e2e"]["value"], "kernels", d["kernels_per_frame"])
print(d["pass_ms"])
PY
tail -3 $OUT/bench_$TAG.err
