#!/bin/bash
# r01k: tile-path rasteriser: parity (raster + mesh-frame tests on both paths) and timing (bench line with the mesh legs).
TAG=${1:-r01k}
OUT=gpurun_out; mkdir -p $OUT
timeout 400 python -m pytest tests/test_raster_gpu.py tests/test_rendergraph_gpu.py -m gpu -x -q > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu_$TAG.log
tail -3 $OUT/pytest_gpu_$TAG.log
LGCU_RASTER_PATH=1 timeout 400 python -m pytest tests/test_raster_gpu.py -m gpu -x -q > $OUT/pytest_oldpath_$TAG.log 2>&1; tail -1 $OUT/pytest_oldpath_$TAG.log
timeout 400 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench exit $?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_r01k.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "from_mesh", d["value_from_mesh"]["value"], d["value_from_mesh"]["ms_per_step"], d["frame_ms"])
print(d["pass_ms_mesh"])
PY
tail -3 $OUT/bench_$TAG.err
