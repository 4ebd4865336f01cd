#!/bin/bash
# Verification of a tree on one B200 (gpurun -- 'bash scripts/gpu_verify.sh TAG'): all GPU tests, smoke(), default bench line, ncu launch
# list, ncu --set full of one frame's kernels and of one mesh frame's raster kernels. Condense with scripts/ncu_summary.py into profiles/.
TAG=${1:-verify}
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu_$TAG.log
tail -4 $OUT/pytest_gpu_$TAG.log
timeout 300 python __graft_entry__.py --smoke > $OUT/smoke_$TAG.log 2>&1; echo "smoke exit $?"; tail -2 $OUT/smoke_$TAG.log
timeout 600 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench exit $?"
cat $OUT/bench_$TAG.json; tail -3 $OUT/bench_$TAG.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > $OUT/ncu_launch_$TAG.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"gatherFast|frameFront|frameChains|chainTail|packDepth|denoiseFinal" -s 12 -c 6 -f -o $OUT/frame_$TAG \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > $OUT/ncu_full_$TAG.log 2>&1
tail -2 $OUT/ncu_full_$TAG.log | cut -c1-200
timeout 400 ncu --set full --clock-control none --import-source on -k regex:raster -s 8 -c 8 -f -o $OUT/raster_$TAG \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > $OUT/ncu_raster_$TAG.log 2>&1
tail -2 $OUT/ncu_raster_$TAG.log | cut -c1-200
ls -la $OUT/*_$TAG.ncu-rep
