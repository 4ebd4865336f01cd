#!/bin/bash
# Runs on the B200 box under gpurun: GPU parity tests, the bench line, the ncu launch list and one full ncu capture of a frame.
# Usage: gpurun --timeout 1500 -- 'bash scripts/gpu_check.sh [tag]'
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi_$TAG.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu_$TAG.log
tail -5 $OUT/pytest_gpu_$TAG.log
timeout 600 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench exit $?"
cat $OUT/bench_$TAG.json
timeout 300 python bench.py --workload 1080p --steps 100 --no-cpu-baseline > $OUT/bench_1080p_$TAG.json 2>> $OUT/bench_$TAG.err
timeout 300 python bench.py --workload 8k --steps 50 --no-cpu-baseline > $OUT/bench_8k_$TAG.json 2>> $OUT/bench_$TAG.err
timeout 300 python bench.py --strict --steps 20 --no-cpu-baseline > $OUT/bench_strict_$TAG.json 2>> $OUT/bench_$TAG.err
timeout 300 python bench.py --mode passes --steps 50 --no-cpu-baseline > $OUT/bench_passes_$TAG.json 2>> $OUT/bench_$TAG.err
# launch list (cold-cache, serialised: shares only)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 1 > $OUT/ncu_launch_$TAG.log 2>&1
# one whole frame of kernels, full sections
timeout 900 ncu --set full --clock-control none --import-source on -s 41 -c 41 -f -o $OUT/frame_$TAG \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 1 > $OUT/ncu_full_$TAG.log 2>&1
ls -la $OUT
