#!/bin/bash
# 2-GPU checks: the C ABI multi-GPU driver (C++ test: NCCL and peer-memory reductions) and bench.py under torchrun at N=2.
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
timeout 600 python -m pytest tests/test_cpp_api.py -q -m gpu -s > gpurun_out/pytest_multi.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_multi.log
tail -6 gpurun_out/pytest_multi.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; echo "bench2 exit $?"
tail -3 gpurun_out/bench_2gpu.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_2gpu.json'))
print('value', d['value']/1e9, 'e2e', d['e2e']['value']/1e9, d['e2e']['host_link_ceiling'], d['e2e']['host_placement'])
print(json.dumps(d['roofline']['secondary']))
print(json.dumps(d.get('capi_multi_join')))
PY
