#!/bin/bash
# strips over N GPUs with the per-rank stage profile: with and without the composite. Usage: bash scripts/gpu_scale4.sh TAG N "workloads"
TAG=${1:-s}; N=${2:-8}; WL=${3:-"8k"}
OUT=gpurun_out; mkdir -p $OUT
run() { timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N "$@"; }
for w in $WL; do
  run --workload $w --shard strips --transport p2p --steps 40 --warmup 5 > $OUT/sc4_strips${w}_n${N}_$TAG.json 2> $OUT/sc4_${w}_n${N}_$TAG.err
  run --workload $w --shard strips --transport p2p --steps 40 --warmup 5 --no-present > $OUT/sc4_strips${w}_nopresent_n${N}_$TAG.json 2> $OUT/sc4_${w}_nopresent_n${N}_$TAG.err
done
for f in $OUT/sc4_*_n${N}_$TAG.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1].split('/')[-1], "ms", round(d["ms_per_step"],3), "Mpix/s", round(d["value"]), "e2e", round(d["e2e"]["value"]))
    for k,v in (d.get("stage_ms_per_rank") or {}).items(): print("   ", k, v)
except Exception as e: print(sys.argv[1], "ERR", e)
PY
done
cat $OUT/sc4_*_n${N}_$TAG.err | grep -v "^\*\|OMP_NUM\|^$" | tail -8
