#!/bin/bash
# r01f: GPU parity tests (incl. mesh-scene frames), default bench line (mesh e2e), gather variants, ncu launch list + full frame capture.
TAG=${1:-r01f}
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu_$TAG.log
tail -4 $OUT/pytest_gpu_$TAG.log
timeout 600 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench exit $?"
cat $OUT/bench_$TAG.json; tail -3 $OUT/bench_$TAG.err
for v in 0 4 7 8 1; do
  LGCU_GATHER_VARIANT=$v timeout 120 python scripts/gather_variants.py >> $OUT/variants_$TAG.jsonl 2>> $OUT/variants_$TAG.err
done
cat $OUT/variants_$TAG.jsonl
for v in 4 7; do
  LGCU_GATHER_VARIANT=$v timeout 300 python -m pytest tests/test_cuda_parity.py tests/test_rendergraph_gpu.py -m gpu -x -q -k "gather or frame or fused" > $OUT/pytest_variant${v}_$TAG.log 2>&1
  echo "variant $v parity: $(tail -1 $OUT/pytest_variant${v}_$TAG.log)"
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > $OUT/ncu_launch_$TAG.log 2>&1
tail -1 $OUT/ncu_launch_$TAG.log | cut -c1-300
