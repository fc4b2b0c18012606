#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_index_gpu.py tests/test_fuzz_gpu.py -x -q -m gpu 2>&1 | tail -3
for m in 1; do S2B_MEPC=$m timeout 300 python tools/time_query.py 2>&1 | tail -1; done
