#!/bin/bash
mkdir -p gpurun_out
for hint in 0 1; do
SEGVLAD_AGG_STORE_HINT=$hint timeout 200 python - <<'PY'
import sys, json, os, torch
sys.path.insert(0, '.')
import bench
peaks, _ = bench._peaks()
for _ in range(2):
    r = bench.aggregation_side_bench(torch.device('cuda'), peaks)
print('store hint', os.environ['SEGVLAD_AGG_STORE_HINT'], json.dumps({k: round(r[k],4) for k in ('ms_per_batch', 'kernel_ms')}), round(r['roofline']['frac'],3))
PY
done
CALLS=3 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_agg.csv python tools/agg_run.py > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/launches_agg.csv')) if len(r)>10]
h=rows[0]; ki=h.index('Kernel Name'); vi=h.index('Metric Value')
seq=[(r[ki].split('(')[0][:60], float(r[vi].replace(',',''))/1000) for r in rows[1:]]
# last call = last third of segvlad kernels
idx=[i for i,(n,_) in enumerate(seq) if 'normalize_centers' in n]
start=idx[-1] if idx else 0
tot=0
for n,us in seq[start:]:
    print(f"{n:60s} {us:9.2f} us"); tot+=us
print('sum', tot)
PY
