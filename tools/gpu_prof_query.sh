#!/bin/bash
mkdir -p gpurun_out
python tools/prof_query.py > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"locate_kernel|process_kernel" -s 6 -c 6 -o gpurun_out/prof_query2 python tools/prof_query.py > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
