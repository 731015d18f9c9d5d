#!/bin/bash
# round-1 re-entry check: GPU parity suite, TC-vs-SIMT aggregation cross-check, bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
for cfg in "2 150 64 8 20" "3 1530 1536 32 150" "2 1530 768 64 200"; do
  set -- $cfg
  B=$1 N=$2 D=$3 K=$4 S=$5 timeout 120 python tools/agg_tc_debug.py 2>&1 | tail -2
done > gpurun_out/agg_tc_debug.log 2>&1
cat gpurun_out/agg_tc_debug.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"
cat gpurun_out/bench_n1.json; tail -3 gpurun_out/bench_n1.err
