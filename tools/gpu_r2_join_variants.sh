#!/bin/bash
# Index-query timings for each experimental build in variants/ and the in-tree library; then the index GPU tests in-tree.
mkdir -p gpurun_out
shopt -s nullglob
for lib in "" variants/*.so; do
  S2B_LIB_PATH=${lib:+$PWD/$lib} timeout 600 python tools/join_time.py 2>&1 | tail -1
done | tee gpurun_out/join_variants.log
timeout 900 python -m pytest tests/test_index_gpu.py tests/test_fuzz_gpu.py -q -m gpu --maxfail=5 2>&1 | tail -5 | tee gpurun_out/pytest_index.log
