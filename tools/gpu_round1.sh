#!/bin/bash
# First GPU pass: tests, smoke, bench, then ncu launch list + full capture of the encode kernel.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
nproc > gpurun_out/nproc.txt; lscpu | head -20 >> gpurun_out/nproc.txt
python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
tail -2 gpurun_out/smoke.log
python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
cat gpurun_out/bench.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err
cat gpurun_out/bench_ref.json | cut -c1-300
python bench.py --steps 3 --warmup 3 --points 20000000 --no-cpu-baseline > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --points 20000000 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
python bench.py --steps 3 --warmup 3 --points 20000000 --no-cpu-baseline > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:encode_kernel -s 3 -c 2 -o gpurun_out/prof_encode python bench.py --steps 3 --warmup 3 --points 20000000 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
