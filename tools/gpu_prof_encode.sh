#!/bin/bash
mkdir -p gpurun_out
python bench.py --steps 3 --warmup 3 --points 20000000 --no-cpu-baseline --no-pip --no-join > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:encode_kernel -s 3 -c 1 -o gpurun_out/prof_encode_r1b python bench.py --steps 3 --warmup 3 --points 20000000 --no-cpu-baseline --no-pip --no-join > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
