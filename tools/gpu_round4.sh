#!/bin/bash
mkdir -p gpurun_out
date +%s > gpurun_out/t0.txt
python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
python -c "
import json
d=json.load(open('gpurun_out/bench.json'))
print('value',d['value'],'e2e',d['e2e']['value'])
for k in ('pip','join','cellunion'): print(k, json.dumps(d.get(k))[:1500])
"
date +%s > gpurun_out/t1.txt
