#!/bin/bash
# ONE frame in row strips over N GPUs (p2p transport, cost-aware bounds, per-rank stage profile).
# Usage: gpurun --gpus N -- 'bash scripts/gpu_strips.sh TAG N "8k 4k"'
TAG=${1:-s}; N=${2:-8}; WL=${3:-"8k"}
OUT=gpurun_out; mkdir -p $OUT
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N "$@"; }
for w in $WL; do
  run --workload $w --shard strips --transport p2p --steps 40 --warmup 5 > $OUT/strips${w}_n${N}_$TAG.json 2> $OUT/strips${w}_n${N}_$TAG.err
  python - "$OUT/strips${w}_n${N}_$TAG.json" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1].split('/')[-1], "ms", round(d["ms_per_step"],3), "Mpix/s", round(d["value"]), "e2e", round(d["e2e"]["value"]), "strips", d["config"]["strips"])
    for k,v in (d.get("stage_ms_per_rank") or {}).items(): print("   ", k, v)
    for b in d.get("balance") or []: print("   balance", b)
except Exception as e: print(sys.argv[1], "ERR", e)
PY
  grep -v "^\*\|OMP_NUM\|^$" $OUT/strips${w}_n${N}_$TAG.err | tail -5
done
