#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; echo "exit $?"
tail -3 gpurun_out/bench_2gpu.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_2gpu.json').read().strip().splitlines()[-1])
print('n_gpus', d['n_gpus'], 'encode', d['value']/1e9, 'ms', d['ms_per_step'], 'e2e', d['e2e']['value']/1e9)
for k in ('pip','join','cellunion'):
    v=d.get(k) or {}
    print(k, {kk: (vv if not isinstance(vv, dict) else '...') for kk, vv in v.items() if kk in ('points_per_s','ms_per_step','total_hits','hits','error','scaling')}, (v.get('uniform') or {}).get('points_per_s'))
PY
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 | cut -c1-200
