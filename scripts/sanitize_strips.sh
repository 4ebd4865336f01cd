#!/bin/bash
# compute-sanitizer over the 2-GPU peer-to-peer strip check: one sanitizer per rank (torchrun's children are not followed reliably).
#   scripts/sanitize_strips.sh memcheck|racecheck TAG
tool=$1; tag=$2
mkdir -p gpurun_out
export MASTER_ADDR=127.0.0.1 MASTER_PORT=29541 WORLD_SIZE=2
for r in 0 1; do
  RANK=$r LOCAL_RANK=$r timeout 1200 compute-sanitizer --tool $tool --error-exitcode 9 --log-file gpurun_out/sanitizer_${tool}_strips_${tag}.rank$r.log \
    python tests/mgpu_strip_check.py 640 384 p2p > gpurun_out/sanitizer_${tool}_strips_${tag}.rank$r.out 2>&1 &
done
wait
for r in 0 1; do tail -1 gpurun_out/sanitizer_${tool}_strips_${tag}.rank$r.out; grep -h "SUMMARY" gpurun_out/sanitizer_${tool}_strips_${tag}.rank$r.log; done
