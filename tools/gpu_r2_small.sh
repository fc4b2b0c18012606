#!/bin/bash
# Targeted GPU tests (cell algebra, C++ API, cell index) — quick confirmation after small changes.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_cellops_gpu.py tests/test_cpp_api.py tests/test_cellindex_gpu.py -q -m gpu --durations=5 2>&1 | tail -25 | tee gpurun_out/pytest_small.log
