#!/bin/bash
# strips with / without graph capture (run under gpurun --gpus N). Usage: bash scripts/gpu_scale2.sh TAG N
TAG=${1:-s}; N=${2:-2}
OUT=gpurun_out; mkdir -p $OUT
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N "$@"; }
run --workload 8k --shard strips --steps 30 --warmup 5 > $OUT/sc2_strips8k_graph_n${N}_$TAG.json 2> $OUT/sc2_n${N}_$TAG.err
run --workload 4k --shard strips --steps 50 --warmup 5 > $OUT/sc2_strips4k_graph_n${N}_$TAG.json 2>> $OUT/sc2_n${N}_$TAG.err
run --workload 512 --shard strips --steps 50 --warmup 5 > $OUT/sc2_strips512_graph_n${N}_$TAG.json 2>> $OUT/sc2_n${N}_$TAG.err
run --workload 512 --shard strips --no-graph --steps 50 --warmup 5 > $OUT/sc2_strips512_nograph_n${N}_$TAG.json 2>> $OUT/sc2_n${N}_$TAG.err
for f in $OUT/sc2_*_n${N}_$TAG.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1].split('/')[-1], "ms", round(d["ms_per_step"],3), "Mpix/s", round(d["value"]), "e2e", round(d["e2e"]["value"]), d["scaling"])
except Exception as e: print(sys.argv[1], "ERR", e)
PY
done
grep -v "^\*\|OMP_NUM\|^$" $OUT/sc2_n${N}_$TAG.err | tail -12
