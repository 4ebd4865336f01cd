#!/bin/bash
# full ncu capture of the gather kernel (with source) at 4K. Usage: gpurun -- 'bash scripts/gpu_ncu_gather.sh TAG'
TAG=${1:-g}
OUT=gpurun_out; mkdir -p $OUT
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gatherFast -s 1 -c 1 -f -o $OUT/gather_$TAG \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 1 > $OUT/ncu_gather_$TAG.log 2>&1
tail -3 $OUT/ncu_gather_$TAG.log
