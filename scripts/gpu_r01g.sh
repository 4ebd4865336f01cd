#!/bin/bash
# r01g: new aux-pass GPU tests, gather tile-size A/B at 4K and on an 8K strip, single-GPU 8K bench line.
TAG=${1:-r01g}
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_aux_passes_gpu.py tests/test_rendergraph_gpu.py tests/test_cuda_parity.py -m gpu -x -q > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu_$TAG.log
tail -4 $OUT/pytest_gpu_$TAG.log
for v in 10 9 0; do
  LGCU_GATHER_VARIANT=$v timeout 120 python scripts/gather_variants.py >> $OUT/variants_$TAG.jsonl 2>> $OUT/variants_$TAG.err
  LGCU_GATHER_VARIANT=$v timeout 120 python scripts/gather_variants.py 7680 4320 2160 2704 >> $OUT/variants_$TAG.jsonl 2>> $OUT/variants_$TAG.err
  LGCU_GATHER_VARIANT=$v timeout 120 python scripts/gather_variants.py 1920 1080 >> $OUT/variants_$TAG.jsonl 2>> $OUT/variants_$TAG.err
done
cat $OUT/variants_$TAG.jsonl; tail -3 $OUT/variants_$TAG.err
timeout 600 python bench.py --workload 8k --steps 50 --warmup 5 --no-cpu-baseline > $OUT/bench_8k_$TAG.json 2> $OUT/bench_8k_$TAG.err; echo "bench 8k exit $?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_8k_r01g.json").read().strip().splitlines()[-1])
print("8k", d["ms_per_step"], d["value"], d["e2e"]["value"], d["value_from_mesh"], d["pass_ms"])
PY
