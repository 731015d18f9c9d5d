#!/bin/bash
# Round evidence: GPU parity suite, smoke, bench line, reference arm, ncu launch list of the bench command, ncu --set full of the
# aggregation / PCA kernels, memcheck over the small cases of the kernels added last.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"; cat gpurun_out/bench_n1.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; cat gpurun_out/bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1; echo "list rc=$?"
if [ -n "$WITH_KNN" ]; then
timeout 900 ncu --set full --import-source on --clock-control none -k regex:'knn_tc_filter|knn_rescore|knn_refine' --launch-skip 21 --launch-count 7 \
  -f -o gpurun_out/knn_final python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-aggregation > gpurun_out/ncu_knn.log 2>&1; echo "ncu knn rc=$?"
fi
CALLS=3 timeout 900 ncu --set full --import-source on --clock-control none -k regex:'aggregate_tc_kernel|rt_from_tokens|assign_tc_kernel' --launch-skip 6 --launch-count 3 \
  -f -o gpurun_out/agg_final python tools/agg_run.py > gpurun_out/ncu_agg.log 2>&1; echo "ncu agg rc=$?"; tail -1 gpurun_out/ncu_agg.log
CALLS=2 timeout 900 ncu --set full --import-source on --clock-control none -k regex:'pca_tc_kernel' --launch-skip 1 --launch-count 1 \
  -f -o gpurun_out/pca_final python tools/pca_run.py > gpurun_out/ncu_pca.log 2>&1; echo "ncu pca rc=$?"; tail -1 gpurun_out/ncu_pca.log
CALLS=3 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_agg.csv python tools/agg_run.py > /dev/null 2>&1
{
echo "compute-sanitizer --tool memcheck (B200, r1 final sources) over small cases of the kernels added in the last session"
echo "--- aggregation (assign_tc, rt_from_tokens, aggregate_tc with bulk tensor stores)"
timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_aggregate.py -q -k "golden or variants or token_major or batched or zero_residual or 64-8" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|Invalid|error" | head -8
echo "--- pca (pca_tc_kernel)"
timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_pca.py -q -k "golden or 129-1000 or 5-64 or 40-512" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|Invalid|error" | head -8
} > gpurun_out/sanitizer.txt 2>&1
cat gpurun_out/sanitizer.txt
ls -la gpurun_out/*.ncu-rep
