#!/bin/bash
mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
python bench.py --steps 2 --warmup 3 --points 20000000 --no-cpu-baseline > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:query_kernel -s 3 -c 4 -o gpurun_out/prof_query python bench.py --steps 2 --warmup 3 --points 20000000 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
