#!/bin/bash
mkdir -p gpurun_out
for cfg in "2 150 64 8 20" "3 1530 1536 32 150"; do
  set -- $cfg
  B=$1 N=$2 D=$3 K=$4 S=$5 timeout 120 python tools/agg_tc_debug.py 2>&1 | tail -1
done
NPASS=26 timeout 120 python tools/agg_tc_timeline.py 2>&1 | tail -28
for pf in 0 1 2 3; do
SEGVLAD_AGG_PREFETCH=$pf timeout 300 python - <<'PY'
import sys, json, torch, os
sys.path.insert(0, '.')
import bench
peaks, _ = bench._peaks()
for _ in range(2):
    r = bench.aggregation_side_bench(torch.device('cuda'), peaks)
print('prefetch', os.environ['SEGVLAD_AGG_PREFETCH'], json.dumps({k: round(r[k], 4) for k in ('ms_per_batch', 'kernel_ms')}), round(r['roofline']['frac'], 3))
PY
done
