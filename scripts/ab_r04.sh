#!/bin/bash
# A/B of the r04 changes (raster tile kernel v2, slim K6+K7, 4-entry pack) on one B200: targeted parity tests, then per-pass times.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "raster or bundled or denoise or final or pack or chained or golden or aux or gi_gather" > gpurun_out/r04a_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r04a_pytest.log
run() { echo "== $*"; env "$@" python scripts/raster_times.py 2>&1 | tail -1; }
{
run A=default
run LGCU_RASTER_RESOLVE_UNROLL=1
run LGCU_RASTER_RESOLVE_UNROLL=4
run LGCU_RASTER_RESOLVE_UNROLL=16
run LGCU_FINAL_ROWS=1
run A=default2
} > gpurun_out/r04a_times.txt 2>&1
cat gpurun_out/r04a_times.txt
