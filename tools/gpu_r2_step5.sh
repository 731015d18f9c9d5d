#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for epi in 8 16; do
SEGVLAD_KNN_EPI=$epi timeout 600 python bench.py --config 4 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --legs none 2>/dev/null | python -c "
import sys, json
for line in sys.stdin:
    if line.startswith('{'):
        d=json.loads(line); print('cfg4 n1 epi=$epi', d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_ms_per_step'])"
SEGVLAD_KNN_EPI=$epi timeout 600 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline --legs none 2>/dev/null | python -c "
import sys, json
for line in sys.stdin:
    if line.startswith('{'):
        d=json.loads(line); print('cfg2 n1 epi=$epi', d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_ms_per_step'])"
done
