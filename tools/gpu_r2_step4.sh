#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_netvlad.py tests/test_gpu_scale.py tests/test_gpu_knn.py -m gpu -x -q 2>&1 | tail -6
timeout 300 python tools/nv_run.py
timeout 600 python bench.py --config 4 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --legs none 2>/dev/null | python -c "
import sys, json
for line in sys.stdin:
    if line.startswith('{'):
        d=json.loads(line); print('cfg4 n1', d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_ms_per_step'])"
