#!/bin/bash
mkdir -p gpurun_out
python tools/prof_query.py > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"locate|process_kernel" -s 9 -c 9 -o gpurun_out/prof_query3 python tools/prof_query.py > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
