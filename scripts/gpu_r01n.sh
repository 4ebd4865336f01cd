#!/bin/bash
# r01n: specialised gather steps (LGCU_GATHER_SPECIALISED) and SFU sRGB encode (LGCU_FAST_SRGB): full parity suite with both ON, A/B timing.
TAG=${1:-r01n}
OUT=gpurun_out; mkdir -p $OUT
LGCU_GATHER_SPECIALISED=1 LGCU_FAST_SRGB=1 timeout 400 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu_on_$TAG.log 2>&1; echo "pytest(on) exit $?" >> $OUT/pytest_gpu_on_$TAG.log
tail -3 $OUT/pytest_gpu_on_$TAG.log
for on in 0 1; do
  LGCU_GATHER_SPECIALISED=$on LGCU_FAST_SRGB=$on timeout 100 python scripts/gather_variants.py >> $OUT/variants_$TAG.jsonl 2>> $OUT/variants_$TAG.err
done
LGCU_GATHER_SPECIALISED=1 timeout 100 python scripts/gather_variants.py 7680 4320 2160 2704 >> $OUT/variants_$TAG.jsonl 2>> $OUT/variants_$TAG.err
cat $OUT/variants_$TAG.jsonl | cut -c1-330; tail -2 $OUT/variants_$TAG.err
LGCU_GATHER_SPECIALISED=1 LGCU_FAST_SRGB=1 timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --e2e-steps 30 > $OUT/bench_on_$TAG.json 2> $OUT/bench_on_$TAG.err; echo "bench exit $?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_on_r01n.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "from_mesh", d["value_from_mesh"]["value"], d["frame_ms"])
print(d["pass_ms"]); print(d["pass_ms_mesh"])
PY
