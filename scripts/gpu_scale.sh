#!/bin/bash
# multi-GPU bench lines (run under gpurun --gpus N). Usage: bash scripts/gpu_scale.sh TAG N
TAG=${1:-s}; N=${2:-2}
OUT=gpurun_out; mkdir -p $OUT
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N "$@"; }
run --workload 8k --shard strips --steps 30 --warmup 5 > $OUT/scale_strips8k_n${N}_$TAG.json 2> $OUT/scale_n${N}_$TAG.err
run --workload 8k --shard strips --no-present --steps 30 --warmup 5 > $OUT/scale_strips8k_nopresent_n${N}_$TAG.json 2>> $OUT/scale_n${N}_$TAG.err
run --workload 4k --shard strips --steps 50 --warmup 5 > $OUT/scale_strips4k_n${N}_$TAG.json 2>> $OUT/scale_n${N}_$TAG.err
run --steps 50 --warmup 5 > $OUT/scale_replicas4k_n${N}_$TAG.json 2>> $OUT/scale_n${N}_$TAG.err
for f in $OUT/scale_*_n${N}_$TAG.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1].split('/')[-1], "ms", round(d["ms_per_step"],3), "Mpix/s", round(d["value"]), "e2e", round(d["e2e"]["value"]), d["scaling"], d["config"].get("exchange_bytes_per_frame_all_ranks"))
except Exception as e: print(sys.argv[1], "ERR", e)
PY
done
tail -5 $OUT/scale_n${N}_$TAG.err
