#!/bin/bash
# r2: full-suite tests + ncu --set full of the non-tensor kernels of the match step (refine, vote, bank prepare, inversion)
# and of the filter kernel at D = 512 (config 4 shape, one GPU)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 900 ncu --set full --import-source on --clock-control none -k regex:'knn_refine|vote_kernel|bank_prepare_kernel|inv_scatter|minmax' --launch-skip 28 --launch-count 9 \
  -f -o gpurun_out/r2_knn_small python bench.py --steps 2 --warmup 3 --no-cpu-baseline --legs none --no-e2e > gpurun_out/ncu_small.log 2>&1; echo "ncu small rc=$?"
timeout 900 ncu --set full --import-source on --clock-control none -k regex:'knn_tc_filter' --launch-skip 10 --launch-count 3 \
  -f -o gpurun_out/r2_filter_d512 python bench.py --config 4 --steps 1 --warmup 3 --no-cpu-baseline --legs none --no-e2e > gpurun_out/ncu_d512.log 2>&1; echo "ncu d512 rc=$?"
ls -la gpurun_out/*.ncu-rep
