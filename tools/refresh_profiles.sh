#!/bin/bash
# Copies the outputs of tools/gpu_evidence.sh from gpurun_out/ into profiles/ (tracked) and rewrites the ncu summaries.
set -e
cp gpurun_out/bench.json profiles/r1_bench_latest.json
cp gpurun_out/bench_ref.json profiles/r1_bench_reference_arm.json
cp gpurun_out/launches.csv profiles/r1_bench_launches.csv
python tools/ncu_summary.py gpurun_out/prof_encode_final.ncu-rep profiles/r1_encode_latlng_final_ncu.json "ncu --set full --clock-control none --import-source on -k regex:encode_kernel -s 3 -c 1 python bench.py --steps 3 --warmup 3 --points 20000000 --join-points 100000000 --no-cpu-baseline --no-pip --no-join" "encode_kernel<latlng>, levels 12/20/30 + ToToken, 20M points (bounded-error sin/cos + reciprocal/rsqrt filter in front of the correctly rounded path). dram traffic per point = (read+write)/20e6." | tail -1
python tools/ncu_summary.py gpurun_out/prof_query_final.ncu-rep profiles/r1_query_final_ncu.json "ncu --set full --clock-control none --import-source on -k regex:'locate|process_kernel|cell_scatter' -s 11 -c 11 python tools/prof_query.py" "K3 at the device granularity (max_edges_per_cell = 1), second (profiled) pass of tools/prof_query.py: U = 20M uniform points vs the 10k-vertex loop (A0 locate_fast, A locate over undecided, B process), near = 5M points in the loop's cap (same three), join = 20M points vs 10k polygons (A0, A, cell_scatter, B over the cell-sorted edge hits, B over interior hits)."
python - <<'PY'
import json
d=json.loads(open('profiles/r1_bench_latest.json').read().strip().splitlines()[-1])
print('encode', round(d['value']/1e9,1), round(d['roofline']['frac'],3), 'e2e', round(d['e2e']['value']/1e9,2), 'xyz', round(d['encode_xyz']['points_per_s']/1e9,1))
p=d['pip']; print('pip', round(p['uniform']['points_per_s']/1e9,1), round(p['near']['points_per_s']/1e9,1), 'default', round(p['default_granularity']['uniform']['points_per_s']/1e9,1), round(p['default_granularity']['near']['points_per_s']/1e9,1))
print('join', round(d['join']['points_per_s']/1e9,1), round(d['join']['roofline']['frac'],3), 'union', round(d['cellunion']['points_per_s']/1e9,1))
print('rj', round(d['radius_join']['count']['targets_per_s']/1e6), round(d['radius_join']['closest']['targets_per_s']/1e6))
PY
