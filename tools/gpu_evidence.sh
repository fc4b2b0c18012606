#!/bin/bash
# Evidence pass: tests, smoke, full bench (both arms), launch list of the bench command (shares, not absolutes), and full
# ncu captures of the dominant kernels. Outputs land in gpurun_out/; summaries are copied into profiles/ by hand.
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu --durations=12 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -16 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>/dev/null
python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
CMD="python bench.py --steps 3 --warmup 3 --points 20000000 --join-points 100000000 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
$CMD --no-pip --no-join > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:encode_kernel -s 3 -c 1 -o gpurun_out/prof_encode_final $CMD --no-pip --no-join > gpurun_out/ncu_full.log 2>&1
python tools/prof_query.py > gpurun_out/plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"locate|process_kernel|cell_scatter" -s 11 -c 11 -o gpurun_out/prof_query_final python tools/prof_query.py > gpurun_out/ncu_full2.log 2>&1
python tools/prof_cellindex.py > gpurun_out/plain4.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:ci_point_histogram -s 3 -c 1 -o gpurun_out/prof_cellindex_final python tools/prof_cellindex.py > gpurun_out/ncu_full3.log 2>&1
ls -la gpurun_out | tail -6
