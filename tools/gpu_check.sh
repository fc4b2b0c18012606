#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench.json'))
print('encode', d['value']/1e9, 'Gpts/s  ms', d['ms_per_step'], 'hbm frac', d['roofline']['frac'], 'e2e', d['e2e']['value']/1e9)
p=d.get('pip') or {}
for k in ('uniform','near'):
    if k in p: print('pip', k, p[k]['points_per_s']/1e9, 'Gpts/s', p[k]['roofline']['frac'])
print('pip e2e', (p.get('e2e_uniform') or {}).get('points_per_s'))
j=d.get('join') or {}; print('join', j.get('points_per_s'), j.get('ms_per_step'), (j.get('cpu_baseline') or {}))
c=d.get('cellunion') or {}; print('union', c.get('points_per_s'), c.get('ms_per_step'), (c.get('cpu_baseline') or {}))
print('cpu', d.get('cpu_baseline'))
PY
