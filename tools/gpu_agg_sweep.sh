#!/bin/bash
for st in 0 15000 30000 45000; do
SEGVLAD_AGG_STAGGER=$st python - <<'PY'
import sys, json, os, torch
sys.path.insert(0, '.')
import bench
peaks, _ = bench._peaks()
for _ in range(2):
    r = bench.aggregation_side_bench(torch.device('cuda'), peaks)
print('stagger', os.environ['SEGVLAD_AGG_STAGGER'], json.dumps({k: round(r[k],4) for k in ('ms_per_batch', 'kernel_ms')}), round(r['roofline']['frac'],3))
PY
done
