#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_encode_gpu.py -x -q -m gpu -s 2>&1 | grep -v "^$" | tail -6
python bench.py --steps 10 --warmup 3 --no-join --no-cpu-baseline 2>&1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
x=d.get('encode_xyz') or {}
print('latlng', round(d['value']/1e9,2), 'frac', round(d['roofline']['frac'],3), 'e2e', d['e2e']['value']/1e9)
print('xyz', x)
p=d.get('pip') or {}
for k in ('uniform','near'):
    if k in p: print('pip', k, p[k]['points_per_s']/1e9, 'Gpts/s', p[k]['roofline']['frac'])
"
