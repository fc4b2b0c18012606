#!/bin/bash
# Launch list of the bench command at HEAD (per-kernel shares; the times under ncu are cold-cache and serialised).
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --points 20000000 --join-points 100000000 --no-cpu-baseline"
timeout 200 $CMD > gpurun_out/plain.log 2>&1 && echo "plain ok" &&
timeout 420 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1; echo "ncu exit $?"
wc -l gpurun_out/launches.csv
