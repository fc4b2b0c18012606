#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | wc -l; free -g | head -2; nproc
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/bench_8gpu.json 2> gpurun_out/bench_8gpu.err; echo "exit $?"
tail -3 gpurun_out/bench_8gpu.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_8gpu.json').read().strip().splitlines()[-1])
print('n_gpus', d['n_gpus'], 'encode', d['value']/1e9, 'ms', d['ms_per_step'], 'e2e', d['e2e'])
for k in ('encode_xyz','pip','join','cellunion'):
    v=d.get(k) or {}
    print(k, {kk: vv for kk, vv in v.items() if kk in ('points_per_s','ms_per_step','total_hits','hits','error','scaling')}, (v.get('uniform') or {}).get('points_per_s'))
PY
