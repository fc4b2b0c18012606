#!/bin/bash
# Bench section of the S2CellIndex join alone (parity on the CPU sample included), no tests / profiler.
mkdir -p gpurun_out
python - > gpurun_out/cellindex_bench.json 2> gpurun_out/cellindex_bench.err <<'PY'
import argparse, json, torch, bench
args = argparse.Namespace(points=100_000_000, steps=10, no_cpu_baseline=False)
print(json.dumps(bench.cell_index_join_section(args, torch, torch.device("cuda:0"), 0, 1)))
PY
echo "bench exit $?"; tail -3 gpurun_out/cellindex_bench.err; cat gpurun_out/cellindex_bench.json
