#!/bin/bash
# Per-kernel durations of the join step (ncu, one metric, no replay of other counters) at bench scale.
mkdir -p gpurun_out
JOIN_POINTS=200000000 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 120 --csv --log-file gpurun_out/join_launches.csv python tools/join_time.py > gpurun_out/join_launches.log 2>&1; echo "ncu exit $?"
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open('gpurun_out/join_launches.csv')) if len(r) > 5]
hdr = rows[0]; ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value'); ui = hdr.index('Metric Unit')
agg = collections.OrderedDict()
for r in rows[1:]:
    name = r[ki].split('(')[0][:70]
    v = float(r[vi].replace(',', '')); 
    if r[ui] == 'ns': v /= 1000.0
    elif r[ui] == 'ms': v *= 1000.0
    agg.setdefault(name, []).append(v)
tot = sum(sum(v) for v in agg.values())
for k, v in agg.items(): print('%-72s n=%3d mean %8.1f us  share %.3f' % (k, len(v), sum(v)/len(v), sum(v)/tot))
PY
