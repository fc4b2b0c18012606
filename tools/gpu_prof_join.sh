#!/bin/bash
mkdir -p gpurun_out
python tools/prof_join.py > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"locate|process_kernel|cell_|DeviceScan" -s 9 -c 9 -o gpurun_out/prof_join5 python tools/prof_join.py > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
