#!/bin/bash
# r01l: branch-free tile rasteriser (parity + timing), gather slice ordering A/B (z-slowest vs x-fastest) with DRAM traffic.
TAG=${1:-r01l}
OUT=gpurun_out; mkdir -p $OUT
timeout 400 python -m pytest tests/test_raster_gpu.py tests/test_rendergraph_gpu.py tests/test_cuda_parity.py -m gpu -x -q > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu_$TAG.log
tail -3 $OUT/pytest_gpu_$TAG.log
for o in 0 1; do
  LGCU_GATHER_ORDER=$o timeout 120 python scripts/gather_variants.py >> $OUT/variants_$TAG.jsonl 2>> $OUT/variants_$TAG.err
  LGCU_GATHER_ORDER=$o timeout 120 python scripts/gather_variants.py 7680 4320 2160 2704 >> $OUT/variants_$TAG.jsonl 2>> $OUT/variants_$TAG.err
  LGCU_GATHER_ORDER=$o timeout 200 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,gpu__time_duration.sum --clock-control none \
      -k regex:gatherFast -s 3 -c 1 --csv --log-file $OUT/gather_traffic_order${o}_$TAG.csv python scripts/gather_variants.py > /dev/null 2>&1
  grep gatherFast $OUT/gather_traffic_order${o}_$TAG.csv | cut -d, -f5,13-
done
cat $OUT/variants_$TAG.jsonl | cut -c1-400
timeout 400 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench exit $?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_r01l.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "from_mesh", d["value_from_mesh"]["value"], d["value_from_mesh"]["ms_per_step"], d["frame_ms"])
print(d["pass_ms"]); print(d["pass_ms_mesh"])
PY
tail -3 $OUT/bench_$TAG.err
