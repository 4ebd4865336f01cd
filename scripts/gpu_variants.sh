#!/bin/bash
# development: compare gather kernel variants (LGCU_GATHER_VARIANT) at 4K. Usage: gpurun -- 'bash scripts/gpu_variants.sh "0 1 2 3"'
OUT=gpurun_out; mkdir -p $OUT
for v in $1; do
  export LGCU_GATHER_VARIANT=$v
  timeout 300 python -m pytest tests -m gpu -x -q -k "gather or golden" > $OUT/pytest_var$v.log 2>&1; echo "variant $v pytest: $(tail -1 $OUT/pytest_var$v.log)"
  timeout 300 python bench.py --steps 50 --no-cpu-baseline > $OUT/bench_var$v.json 2>$OUT/bench_var$v.err
  python - <<PY
import json
d=json.loads(open("$OUT/bench_var$v.json").read().strip().splitlines()[-1])
print("variant $v: frame ms", round(d["ms_per_step"],4), "gather ms", d["pass_ms"].get("IndirectLightPass"))
PY
done
