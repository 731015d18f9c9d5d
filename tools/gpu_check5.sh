#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_knn.py tests/test_gpu_e2e.py tests/test_gpu_scale.py -q -x > gpurun_out/pytest_knn.log 2>&1; echo "knn rc=$?"
tail -5 gpurun_out/pytest_knn.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-aggregation > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench_n1.json')); r=d['roofline']; print('ms/step', round(d['ms_per_step'],3), 'tc_ms', round(r['kernel_ms_per_step'],3), 'rescore', round(r['rescore_ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), 'value', d['value']/1e9, 'e2e', d['e2e']['value']/1e9)"
tail -3 gpurun_out/bench_n1.err
