#!/bin/bash
# 8-GPU (and 4-GPU) bench lines at HEAD: torchrun, one rank per GPU; rank 0 also drives all GPUs through the C ABI.
mkdir -p gpurun_out
nvidia-smi -L | wc -l; free -g | head -2; nproc
for N in 8 4; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err; echo "N=$N exit $?"
  tail -2 gpurun_out/bench_${N}gpu.err
  python - $N <<'PY'
import json, sys
d=json.load(open('gpurun_out/bench_%sgpu.json' % sys.argv[1]))
print('n_gpus', d['n_gpus'], 'encode', d['value']/1e9, 'ms', d['ms_per_step'], 'e2e', d['e2e']['value']/1e9, d['e2e']['host_link_ceiling'], d['e2e']['host_placement'])
print(json.dumps(d['roofline']['secondary']))
print(json.dumps(d.get('capi_multi_join')))
PY
done
