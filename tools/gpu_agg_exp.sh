#!/bin/bash
for mb in 0 32 64 96; do
SEGVLAD_AGG_L2_VERBOSE=1 SEGVLAD_AGG_L2_MB=$mb timeout 200 python - <<'PY'
import sys, json, os, torch
sys.path.insert(0, '.')
import bench
peaks, _ = bench._peaks()
for _ in range(2):
    r = bench.aggregation_side_bench(torch.device('cuda'), peaks)
print('L2 MB', os.environ['SEGVLAD_AGG_L2_MB'], json.dumps({k: round(r[k],4) for k in ('ms_per_batch', 'kernel_ms')}))
PY
done
