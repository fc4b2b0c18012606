#!/bin/bash
# ncu --set full of the S2CellIndex point join kernel (100M points, the bench scene).
mkdir -p gpurun_out
timeout 300 python tools/prof_cellindex.py 100000000 > gpurun_out/prof_cellindex_plain.log 2>&1 && echo plain ok &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ci_point_histogram -s 3 -c 1 -f -o gpurun_out/cellindex_r2 python tools/prof_cellindex.py 100000000 > gpurun_out/ncu_cellindex.log 2>&1; echo "ncu exit $?"
