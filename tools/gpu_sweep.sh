#!/bin/bash
mkdir -p gpurun_out
for f in 4 16 64 256; do S2B_LUT_FACTOR=$f python tools/time_query.py 2>&1 | tail -1; done | tee gpurun_out/sweep.txt
S2B_LUT_FACTOR=64 python -c "
import ctypes
from s2geometry_b200 import _lib
_lib.load().s2b_diag_force_exact_locate(1)
import runpy, sys
sys.argv=['x']
runpy.run_path('tools/time_query.py', run_name='__main__')
" 2>&1 | tail -1
