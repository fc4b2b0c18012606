#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_encode_gpu.py tests/test_index_gpu.py -x -q -m gpu 2>&1 | tail -4
python tools/time_encode_xyz.py | tail -1
S2B_NO_STAGING=1 python tools/time_encode_xyz.py | tail -1
