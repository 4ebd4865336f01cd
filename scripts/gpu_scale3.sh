#!/bin/bash
# strips over N GPUs, p2p vs nccl transport (run under gpurun --gpus N). Usage: bash scripts/gpu_scale3.sh TAG N "workloads" "transports"
TAG=${1:-s}; N=${2:-2}; WL=${3:-"512 4k 8k"}; TR=${4:-"p2p nccl"}
OUT=gpurun_out; mkdir -p $OUT
run() { timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N "$@"; }
for w in $WL; do for t in $TR; do
  run --workload $w --shard strips --transport $t --steps 40 --warmup 5 > $OUT/sc3_strips${w}_${t}_n${N}_$TAG.json 2> $OUT/sc3_${w}_${t}_n${N}_$TAG.err
done; done
for f in $OUT/sc3_*_n${N}_$TAG.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1].split('/')[-1], "ms", round(d["ms_per_step"],3), "Mpix/s", round(d["value"]), "e2e", round(d["e2e"]["value"]), d["scaling"], "clk", d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
except Exception as e: print(sys.argv[1], "ERR", e)
PY
done
cat $OUT/sc3_*_n${N}_$TAG.err | grep -v "^\*\|OMP_NUM\|^$" | tail -8
