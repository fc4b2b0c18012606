#!/bin/bash
# Strict-path tests, a quick headline bench, then ONE ncu --set full capture (with SASS source counters) of the headline
# encode kernel at the bench's own size. Outputs under gpurun_out/.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_encode_gpu.py tests/test_cpp_api.py -q -m gpu --maxfail=5 > gpurun_out/pytest_encode.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_encode.log
tail -8 gpurun_out/pytest_encode.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-pip --no-join --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "bench exit $?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_quick.json'))
print('value', d['value']/1e9, 'ms', d['ms_per_step'], 'kernel_ms', d['roofline']['kernel_ms'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['value']/1e9, 'launches', d['gpu_launches'])
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:encode_kernel -s 3 -c 1 -f -o gpurun_out/encode_r2 \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-pip --no-join > gpurun_out/ncu_encode.log 2>&1; echo "ncu exit $?"
ls -la gpurun_out/*.ncu-rep
