#!/bin/bash
# Point-index GPU tests.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_pointindex_gpu.py -q -m gpu -x --durations=5 2>&1 | tail -30 | tee gpurun_out/pytest_pi.log
