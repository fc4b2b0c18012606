#!/bin/bash
# bench.py must print exactly ONE line on stdout (the JSON), at N=1 and under torchrun.
mkdir -p gpurun_out
F="--steps 3 --warmup 3 --no-pip --no-join --no-cpu-baseline"
python bench.py $F > gpurun_out/so1.txt 2>/dev/null; echo "n1 exit $? lines $(wc -l < gpurun_out/so1.txt)"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 $F > gpurun_out/so2.txt 2>/dev/null; echo "n2 exit $? lines $(wc -l < gpurun_out/so2.txt)"
head -c 120 gpurun_out/so1.txt; echo; head -c 120 gpurun_out/so2.txt; echo
python -c "import json;[json.loads(open(f).read()) for f in ('gpurun_out/so1.txt','gpurun_out/so2.txt')];print('both parse as one JSON document')"
