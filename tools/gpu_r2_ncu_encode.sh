#!/bin/bash
# ONE ncu --set full capture (with SASS source counters) of the headline encode kernel at the bench's own size.
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:encode_kernel -s 3 -c 1 -f -o gpurun_out/encode_r2 \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-pip --no-join > gpurun_out/ncu_encode.log 2>&1; echo "ncu exit $?"
ls -la gpurun_out/*.ncu-rep
