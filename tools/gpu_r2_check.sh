#!/bin/bash
# Round-2 GPU check: full GPU test suite, the bench line (N=1), the reference arm. Outputs under gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt; free -g | head -2 >> gpurun_out/gpu.txt
timeout 1500 python -m pytest tests -q -m gpu --maxfail=8 --durations=15 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -30 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
tail -5 gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref exit $?"
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/bench.json'))
    print(json.dumps({k: d[k] for k in ('value','ms_per_step','gpu_launches','clocks')}))
    print('roofline', {k: v for k, v in d['roofline'].items() if k != 'fp64'})
    print('e2e', d['e2e'])
    print('cpu', d.get('cpu_baseline'))
    print('summary', json.dumps(d['roofline']['secondary'], indent=0))
    r=json.load(open('gpurun_out/bench_ref.json')); print('ref', r['value'], r['ms_per_step'], r['cpu_baseline'])
except Exception as e:
    print('summary failed', e)
PY
