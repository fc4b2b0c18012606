#!/bin/bash
# First GPU pass of the S2CellIndex joins: parity tests, then the bench section alone (short encode leg, no PIP / polygon join).
mkdir -p gpurun_out
python -m pytest tests/test_cellindex_gpu.py -x -q > gpurun_out/pytest_cellindex.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_cellindex.log
tail -15 gpurun_out/pytest_cellindex.log
python - > gpurun_out/cellindex_bench.json 2> gpurun_out/cellindex_bench.err <<'PY'
import argparse, json, torch, bench
args = argparse.Namespace(points=100_000_000, steps=10, no_cpu_baseline=False)
dev = torch.device("cuda:0")
print(json.dumps(bench.cell_index_join_section(args, torch, dev, 0, 1)))
PY
echo "bench exit $?"; tail -3 gpurun_out/cellindex_bench.err; cat gpurun_out/cellindex_bench.json
python tools/prof_cellindex.py > gpurun_out/prof_cellindex_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:ci_point_histogram -s 3 -c 1 -o gpurun_out/prof_cellindex python tools/prof_cellindex.py > gpurun_out/ncu_cellindex.log 2>&1
tail -2 gpurun_out/ncu_cellindex.log
