#!/bin/bash
for e in 0 1 16 17; do
SEGVLAD_AGG_EXP=$e timeout 200 python - <<'PY'
import sys, json, os, torch
sys.path.insert(0, '.')
import bench
peaks, _ = bench._peaks()
for _ in range(2):
    r = bench.aggregation_side_bench(torch.device('cuda'), peaks)
print('exp', os.environ['SEGVLAD_AGG_EXP'], json.dumps({k: round(r[k],4) for k in ('ms_per_batch', 'kernel_ms')}))
PY
done
