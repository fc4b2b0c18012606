#!/bin/bash
# Evidence pass: launch list of the bench command (shares, not absolutes) + full captures of the dominant kernels.
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --points 20000000 --join-points 100000000 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
$CMD --no-pip --no-join > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:encode_kernel -s 3 -c 1 -o gpurun_out/prof_encode3 $CMD --no-pip --no-join > gpurun_out/ncu_full.log 2>&1
python tools/prof_join.py > gpurun_out/plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"locate_kernel|process_kernel" -s 3 -c 3 -o gpurun_out/prof_join3 python tools/prof_join.py > gpurun_out/ncu_full2.log 2>&1
ls -la gpurun_out | tail -8
