#!/bin/bash
# single-sweep ("resident") aggregation: parity tests, then the bench's aggregation / fused legs with the mode on and off
mkdir -p gpurun_out
timeout 150 python -m pytest tests/test_gpu_aggregate.py tests/test_gpu_pca.py tests/test_gpu_e2e.py tests/test_gpu_anyloc.py -m gpu -x -q 2>&1 | tail -8
if [ ${PIPESTATUS[0]} -ne 0 ]; then echo "tests failed: stopping"; exit 1; fi
run() {
timeout 90 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --legs aggregation,pca 2>gpurun_out/res_$1.err | python -c "
import sys, json
for line in sys.stdin:
    if line.startswith('{'):
        d=json.loads(line); a=d['aggregation']; f=d.get('aggregate_pca_fused') or {}
        print('$1', 'kernel_ms', a['kernel_ms'], 'frac', a['roofline']['frac'], 'batch_ms', a['ms_per_batch'], 'fused', f.get('fused'))"
}
run default
SEGVLAD_AGG_LA=1 run lookahead
SEGVLAD_AGG_RESIDENT=0 run res0
tail -3 gpurun_out/res_default.err
