#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_pca.py tests/test_gpu_e2e.py -q -x > gpurun_out/pytest_pca.log 2>&1; echo "pca rc=$?"; tail -3 gpurun_out/pytest_pca.log
timeout 300 python - <<'PY'
import sys, json, torch, os
sys.path.insert(0, '.')
import bench
peaks, _ = bench._peaks()
r = bench.pca_side_bench(torch.device('cuda'), peaks)
print('tc  ', json.dumps({k: (round(v, 4) if isinstance(v, float) else v) for k, v in r.items() if k != 'roofline'}), round(r['roofline']['achieved'], 1), 'TFLOP/s algorithmic')
PY
