#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_knn.py -q -x > gpurun_out/pytest_knn.log 2>&1; echo "knn rc=$?"
tail -3 gpurun_out/pytest_knn.log
for cfg in "100000 6" "100000 8" "100000 100" "48 6" "48 8" "96 8" "24 8"; do
  set -- $cfg
  SEGVLAD_RESCORE_WAVE_MB=$1 SEGVLAD_RESCORE_CTAS=$2 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-aggregation 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('wave_mb=$1 ctas=$2', 'ms/step', round(d['ms_per_step'],3), 'tc_ms', round(r['kernel_ms_per_step'],3), 'rescore', round(r['rescore_ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3))"
done
