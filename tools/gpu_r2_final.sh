#!/bin/bash
# Round-2 evidence run on one B200: GPU test suite, smoke(), bench (N=1), reference arm, then (each only after the plain
# command exited 0) the launch list of the bench command and one ncu --set full capture of the encode and join kernels.
mkdir -p gpurun_out
bash tools/gpu_r2_check.sh
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/smoke.log
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
timeout 300 $CMD > gpurun_out/plain.json 2> gpurun_out/plain.err && echo "plain ok" &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1; echo "launch list exit $?"
wc -l gpurun_out/launches.csv
bash tools/gpu_r2_ncu_encode.sh
bash tools/gpu_r2_ncu_join.sh
