#!/bin/bash
# ncu --set full (with source counters) of one launch each of the join's kernels at bench scale (64M-point round).
mkdir -p gpurun_out
JOIN_POINTS=128000000 timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'locate_fast_kernel|locate_kernel|process_kernel|cell_scatter' -s 24 -c 6 -f -o gpurun_out/join_r2 python tools/join_time.py > gpurun_out/ncu_join.log 2>&1; echo "ncu exit $?"
ls -la gpurun_out/join_r2.ncu-rep
