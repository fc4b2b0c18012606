#!/bin/bash
mkdir -p gpurun_out
python tools/prof_join.py > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"locate_kernel|process_kernel" -s 3 -c 3 -o gpurun_out/prof_join2 python tools/prof_join.py > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
