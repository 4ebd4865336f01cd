#!/bin/bash
# r01h: re-run of the parity tests after the MipLevelArgs fix, gather pattern-slice A/B (work-unit granularity).
TAG=${1:-r01h}
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_cuda_parity.py tests/test_rendergraph_gpu.py -m gpu -x -q > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu_$TAG.log
tail -3 $OUT/pytest_gpu_$TAG.log
for cfg in "10 1" "10 2" "10 4" "10 16" "9 4" "9 16"; do
  set -- $cfg
  LGCU_GATHER_VARIANT=$1 LGCU_GATHER_SLICES=$2 timeout 120 python scripts/gather_variants.py >> $OUT/variants_$TAG.jsonl 2>> $OUT/variants_$TAG.err
  LGCU_GATHER_VARIANT=$1 LGCU_GATHER_SLICES=$2 timeout 120 python scripts/gather_variants.py 7680 4320 2160 2704 >> $OUT/variants_$TAG.jsonl 2>> $OUT/variants_$TAG.err
done
cat $OUT/variants_$TAG.jsonl; tail -3 $OUT/variants_$TAG.err
LGCU_GATHER_VARIANT=10 LGCU_GATHER_SLICES=4 timeout 300 python -m pytest tests/test_cuda_parity.py -m gpu -x -q -k "gather" > $OUT/pytest_slices_$TAG.log 2>&1; tail -1 $OUT/pytest_slices_$TAG.log
