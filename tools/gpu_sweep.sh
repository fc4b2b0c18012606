#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_index_gpu.py -x -q -m gpu 2>&1 | tail -2
for v in 0 1 2 3 4; do S2B_LOCATE_VARIANT=$v python tools/time_query.py 2>&1 | tail -1; done | tee gpurun_out/sweep.txt
