#!/bin/bash
# Region cell tests: GPU tests (region, locate, C++ API incl. replicas) + the bench section alone.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_region_gpu.py tests/test_index_gpu.py tests/test_cpp_api.py tests/test_pointindex_gpu.py -q -m gpu 2>&1 | tail -3
timeout 600 python tools/region_bench.py 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('may_intersect', d['may_intersect']['targets_per_s']/1e9, 'contains', d['contains_cell']['targets_per_s']/1e9, 'parity', d.get('cpu_baseline',{}).get('parity_on_sample'))
"
