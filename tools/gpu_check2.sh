#!/bin/bash
# kNN v2 (single fp16 pass) check: kNN parity tests, full GPU suite, bench, ncu launch list
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_knn.py -q -x > gpurun_out/pytest_knn.log 2>&1; echo "knn rc=$?"
tail -15 gpurun_out/pytest_knn.log
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "all rc=$?"
tail -8 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"
cat gpurun_out/bench_n1.json; tail -3 gpurun_out/bench_n1.err
SEGVLAD_KNN_CTAS=1 timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-aggregation > gpurun_out/bench_n1_cta1.json 2>&1
cat gpurun_out/bench_n1_cta1.json
