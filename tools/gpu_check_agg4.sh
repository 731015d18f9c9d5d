#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_aggregate.py tests/test_gpu_anyloc.py tests/test_gpu_e2e.py tests/test_gpu_pca.py -q -x > gpurun_out/pytest_agg.log 2>&1; echo "agg rc=$?"; tail -15 gpurun_out/pytest_agg.log
bash tools/gpu_agg_list.sh 2>&1 | tail -14
