#!/bin/bash
# full ncu capture of one fused frame (all kernels, with source) at 4K + the launch list. Usage: gpurun -- 'bash scripts/gpu_ncu_frame.sh TAG [kernels_per_frame]'
TAG=${1:-f}; N=${2:-5}
OUT=gpurun_out; mkdir -p $OUT
timeout 900 ncu --set full --clock-control none --import-source on -s $N -c $N -f -o $OUT/frame_$TAG \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 1 > $OUT/ncu_frame_$TAG.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 1 > $OUT/ncu_launch_$TAG.log 2>&1
tail -2 $OUT/ncu_frame_$TAG.log
