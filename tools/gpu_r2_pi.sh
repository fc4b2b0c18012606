#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_pointindex_gpu.py tests/test_region_gpu.py tests/test_index_gpu.py -q -m gpu --maxfail=5 2>&1 | tail -30 | tee gpurun_out/pytest_pi.log
