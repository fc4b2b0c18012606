#!/bin/bash
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 3 --points 20000000 --join-points 100000000 --no-cpu-baseline --no-pip > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"query_kernel|cellunion" -s 3 -c 4 -o gpurun_out/prof_join python bench.py --steps 2 --warmup 3 --points 20000000 --join-points 100000000 --no-cpu-baseline --no-pip > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
