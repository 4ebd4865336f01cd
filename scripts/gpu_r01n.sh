#!/bin/bash
# r01n: switch-guarded optimisations — specialised gather steps (LGCU_GATHER_SPECIALISED), SFU sRGB encode (LGCU_FAST_SRGB), frame-front
# occupancy target (LGCU_FRONT_BLOCKS): full parity suite with all ON, A/B pass timings, bench line with all ON.
TAG=${1:-r01n}
OUT=gpurun_out; mkdir -p $OUT
export LGCU_GATHER_SPECIALISED=1 LGCU_FAST_SRGB=1 LGCU_FRONT_BLOCKS=3
timeout 400 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu_on_$TAG.log 2>&1; echo "pytest(on) exit $?" >> $OUT/pytest_gpu_on_$TAG.log
tail -3 $OUT/pytest_gpu_on_$TAG.log
LGCU_FRONT_BLOCKS=4 timeout 200 python -m pytest tests/test_cuda_parity.py tests/test_rendergraph_gpu.py -m gpu -x -q -k "front or frame or fused" > $OUT/pytest_front4_$TAG.log 2>&1; echo "front4: $(tail -1 $OUT/pytest_front4_$TAG.log)"
timeout 100 python scripts/gather_variants.py >> $OUT/variants_$TAG.jsonl 2>> $OUT/variants_$TAG.err
LGCU_FRONT_BLOCKS=4 timeout 100 python scripts/gather_variants.py >> $OUT/variants_$TAG.jsonl 2>> $OUT/variants_$TAG.err
LGCU_GATHER_SPECIALISED=0 LGCU_FAST_SRGB=0 LGCU_FRONT_BLOCKS=2 timeout 100 python scripts/gather_variants.py >> $OUT/variants_$TAG.jsonl 2>> $OUT/variants_$TAG.err
cat $OUT/variants_$TAG.jsonl | cut -c1-360; tail -2 $OUT/variants_$TAG.err
timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --e2e-steps 30 > $OUT/bench_on_$TAG.json 2> $OUT/bench_on_$TAG.err; echo "bench exit $?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_on_r01n.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "from_mesh", d["value_from_mesh"]["value"], d["frame_ms"])
print(d["pass_ms"]); print(d["pass_ms_mesh"])
PY
