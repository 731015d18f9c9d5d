#!/bin/bash
# Round-2 evidence: GPU parity suite, smoke, bench line (all legs), reference arm, ncu launch list of the bench command,
# ncu --set full of the kernels added / changed this round, memcheck over their small cases.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err; echo "ref rc=$?"; cut -c1-400 gpurun_out/r2_bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_bench_n1.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --legs aggregation,pca --no-e2e > gpurun_out/ncu_list.log 2>&1; echo "list rc=$?"
timeout 900 ncu --set full --import-source on --clock-control none -k regex:'knn_rescore_ref|knn_tc_filter|merge_packed' --launch-skip 9 --launch-count 4 \
  -f -o gpurun_out/r2_knn_final python bench.py --steps 2 --warmup 3 --no-cpu-baseline --legs none --no-e2e > gpurun_out/ncu_knn.log 2>&1; echo "ncu knn rc=$?"
CALLS=2 timeout 900 ncu --set full --import-source on --clock-control none -k regex:'aggregate_tc_kernel|pca_tc_kernel' --launch-skip 2 --launch-count 2 \
  -f -o gpurun_out/r2_fused python tools/fused_run.py > gpurun_out/ncu_fused.log 2>&1; echo "ncu fused rc=$?"
timeout 900 ncu --set full --import-source on --clock-control none -k regex:'nv_assign_tc|nv_vlad_tc|nv_prep' --launch-skip 3 --launch-count 3 \
  -f -o gpurun_out/r2_netvlad python tools/nv_run.py > gpurun_out/ncu_nv.log 2>&1; echo "ncu nv rc=$?"
{
echo "compute-sanitizer --tool memcheck (B200, r2 final sources) over small cases of the kernels added / changed in round 2"
timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_knn.py tests/test_gpu_vote.py -q -k "small_vs_oracle or async or merge_packed or fewer_refs or partial_seg" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|Invalid|error" | head -8
timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_pca.py tests/test_gpu_netvlad.py tests/test_gpu_aggregate.py -q -k "fused or netvlad or skewed or tiles_at_image or mask_centroids" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|Invalid|error" | head -8
} > gpurun_out/r2_sanitizer_memcheck.txt 2>&1
cat gpurun_out/r2_sanitizer_memcheck.txt
ls -la gpurun_out/r2_*.ncu-rep
