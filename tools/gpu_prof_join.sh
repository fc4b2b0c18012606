#!/bin/bash
mkdir -p gpurun_out
python tools/time_query.py 2>&1 | tail -1
python tools/prof_join.py > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"locate|process_kernel" -s 4 -c 4 -o gpurun_out/prof_join4 python tools/prof_join.py > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
