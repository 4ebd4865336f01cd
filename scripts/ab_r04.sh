#!/bin/bash
# A/B of the r04b changes (tail + blur grid in one cluster launch, L1 prefetch in the frame front) on one B200, and an ncu --set full
# capture of the kernels this round touched.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "chain or front or golden or chained or rendergraph or strips or full_size" > gpurun_out/r04b_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r04b_pytest.log
run() { echo "== $*"; env "$@" python scripts/raster_times.py 2>&1 | tail -1; }
{
run A=default
run LGCU_CHAINS_SPLIT=1
run LGCU_FRONT_PREFETCH=0
run A=default2
} > gpurun_out/r04b_times.txt 2>&1
cat gpurun_out/r04b_times.txt
timeout 300 ncu --set full --import-source on --clock-control none -k regex:"rasterTileKernel|denoiseFinalGather|packSidePyramid|frameFront|frameChains" -c 14 -o gpurun_out/r04b_small_kernels python scripts/raster_times.py > gpurun_out/r04b_ncu.log 2>&1
echo "ncu rc=$?"; ls -la gpurun_out/r04b_small_kernels.ncu-rep
