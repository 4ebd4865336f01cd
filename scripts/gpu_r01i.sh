#!/bin/bash
# r01i (2 GPUs): all GPU tests incl. the multi-GPU ones, then the strips bench with cost-aware bounds at N GPUs.
TAG=${1:-r01i}; N=${2:-2}
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu_$TAG.log
tail -4 $OUT/pytest_gpu_$TAG.log
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N "$@"; }
run --workload 8k --shard strips --transport p2p --steps 40 --warmup 5 > $OUT/strips8k_n${N}_$TAG.json 2> $OUT/strips8k_n${N}_$TAG.err
python - "$OUT/strips8k_n${N}_$TAG.json" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1].split('/')[-1], "ms", round(d["ms_per_step"],3), "Mpix/s", round(d["value"]), "e2e", round(d["e2e"]["value"]), "strips", d["config"]["strips"])
    for k,v in (d.get("stage_ms_per_rank") or {}).items(): print("   ", k, v)
    for b in d.get("balance") or []: print("   balance", b)
except Exception as e: print(sys.argv[1], "ERR", e)
PY
grep -v "^\*\|OMP_NUM\|^$" $OUT/strips8k_n${N}_$TAG.err | tail -8
