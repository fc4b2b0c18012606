#!/bin/bash
mkdir -p gpurun_out
for v in 2 3 4; do S2B_PROCESS_VARIANT=$v python tools/time_query.py 2>&1 | tail -1; done | tee gpurun_out/sweep.txt
