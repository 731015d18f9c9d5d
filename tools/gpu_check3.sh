#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_knn.py tests/test_gpu_e2e.py tests/test_gpu_scale.py -q -x > gpurun_out/pytest_knn.log 2>&1; echo "knn rc=$?"
tail -15 gpurun_out/pytest_knn.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-aggregation > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"
cat gpurun_out/bench_n1.json; tail -3 gpurun_out/bench_n1.err
timeout 900 ncu --set full --import-source on --clock-control none -k regex:knn_tc_filter --launch-skip 9 --launch-count 3 \
  -f -o gpurun_out/knn_v3 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-aggregation > gpurun_out/ncu_knn.log 2>&1
echo "ncu rc=$?"
