#!/bin/bash
# Quick check of the lat/lng path: its GPU tests + the headline bench without the secondary sections.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_encode_gpu.py tests/test_cpp_api.py -q -m gpu --maxfail=5 > gpurun_out/pytest_encode.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_encode.log
tail -8 gpurun_out/pytest_encode.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-pip --no-join --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "bench exit $?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_quick.json'))
print('value', d['value']/1e9, 'ms', d['ms_per_step'], 'kernel_ms', d['roofline']['kernel_ms'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['value']/1e9, 'launches', d['gpu_launches'], d['roofline']['secondary']['strict_reevaluated_per_step'])
PY
