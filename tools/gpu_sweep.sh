#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_encode_gpu.py -x -q -m gpu -s 2>&1 | grep -v "^$" | tail -6
python bench.py --steps 10 --warmup 3 --no-pip --no-join --no-cpu-baseline 2>&1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('encode', d['value']/1e9, 'frac', d['roofline']['frac'], 'e2e', d['e2e']['value']/1e9, {k:v for k,v in d.items() if k in ('encode_xyz',)})
"
