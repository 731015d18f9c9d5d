#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/agg_tc_probe.py 2>&1 | tail -20
timeout 900 python -m pytest tests/test_gpu_aggregate.py tests/test_gpu_anyloc.py tests/test_gpu_e2e.py tests/test_gpu_pca.py -q -x > gpurun_out/pytest_agg.log 2>&1; echo "agg rc=$?"; tail -4 gpurun_out/pytest_agg.log
for cfg in "2 150 64 8 20" "3 1530 1536 32 150" "2 1530 768 64 200" "1 1530 1536 32 300"; do
  set -- $cfg
  B=$1 N=$2 D=$3 K=$4 S=$5 timeout 120 python tools/agg_tc_debug.py 2>&1 | tail -1
done
timeout 300 python - <<'PY'
import sys, json, torch
sys.path.insert(0, '.')
import bench
peaks, _ = bench._peaks()
for _ in range(2):
    r = bench.aggregation_side_bench(torch.device('cuda'), peaks)
print(json.dumps({k: r[k] for k in ('ms_per_batch', 'kernel_ms')}), r['roofline']['frac'])
PY
