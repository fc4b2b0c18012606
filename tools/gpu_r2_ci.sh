#!/bin/bash
# S2CellIndex join: GPU tests + the bench section alone.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_cellindex_gpu.py -q -m gpu 2>&1 | tail -2
bash tools/gpu_cellindex_bench.sh 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print('cell_index_join', d['points_per_s']/1e9, 'Gpts/s', d['roofline']['frac'], 'parity', d.get('cpu_baseline',{}).get('parity_on_sample'), 'sparse', (d.get('sparse') or {}).get('points_per_s'))
    elif 'exit' in l: print(l)
"
