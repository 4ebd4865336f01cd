#!/bin/bash
# r01e: re-baseline after container re-creation: pipe-rate microbench, GPU parity tests, default bench line, full ncu capture of the gather.
TAG=${1:-r01e}
OUT=gpurun_out; mkdir -p $OUT
./scripts/microbench/ffma2 > $OUT/ffma2_$TAG.txt 2>&1; cat $OUT/ffma2_$TAG.txt
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu_$TAG.log
tail -4 $OUT/pytest_gpu_$TAG.log
timeout 600 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench exit $?"
cat $OUT/bench_$TAG.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gatherFast -s 1 -c 1 -f -o $OUT/gather_$TAG \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 1 > $OUT/ncu_gather_$TAG.log 2>&1
tail -2 $OUT/ncu_gather_$TAG.log
