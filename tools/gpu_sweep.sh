#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_index_gpu.py tests/test_fuzz_gpu.py tests/test_cpp_api.py -x -q -m gpu 2>&1 | tail -3
S2B_SORT_BY_CELL=1 python -m pytest tests/test_index_gpu.py tests/test_fuzz_gpu.py -x -q -m gpu 2>&1 | tail -2
python tools/time_query.py 2>&1 | tail -1
