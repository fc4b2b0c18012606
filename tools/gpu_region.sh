#!/bin/bash
# Region cell tests on the GPU: the bench section alone, then a full ncu capture of the two MayIntersect kernels.
mkdir -p gpurun_out
python tools/region_bench.py --no-cpu-baseline > gpurun_out/plain_region.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"region_classify|region_process" -s 6 -c 2 -o gpurun_out/prof_region python tools/region_bench.py --no-cpu-baseline > gpurun_out/ncu_region.log 2>&1
ls -la gpurun_out | tail -4
