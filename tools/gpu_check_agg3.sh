#!/bin/bash
mkdir -p gpurun_out
for cfg in "2 150 64 8 20" "3 1530 1536 32 150" "2 1530 768 64 200" "1 1530 1536 32 300"; do
  set -- $cfg
  B=$1 N=$2 D=$3 K=$4 S=$5 timeout 120 python tools/agg_tc_debug.py 2>&1 | tail -1
done
timeout 900 python -m pytest tests/test_gpu_aggregate.py tests/test_gpu_anyloc.py tests/test_gpu_e2e.py tests/test_gpu_pca.py -q -x > gpurun_out/pytest_agg.log 2>&1; echo "agg rc=$?"; tail -4 gpurun_out/pytest_agg.log
timeout 300 python tools/agg_tc_probe.py 2>&1 | tail -20
for e in 0 1 8; do
SEGVLAD_AGG_EXP=$e timeout 200 python - <<'PY'
import sys, json, os, torch
sys.path.insert(0, '.')
import bench
peaks, _ = bench._peaks()
for _ in range(2):
    r = bench.aggregation_side_bench(torch.device('cuda'), peaks)
print('exp', os.environ['SEGVLAD_AGG_EXP'], json.dumps({k: round(r[k],4) for k in ('ms_per_batch', 'kernel_ms')}), round(r['roofline']['frac'],3))
PY
done
python - <<'PY'
import torch
x = torch.empty(1610612736 // 8, dtype=torch.float64, device='cuda')
for _ in range(3): x.zero_()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): x.zero_()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print('pure write 1.61 GB: %.4f ms = %.0f GB/s' % (ms, 1.610612736 / ms * 1e3))
PY
