#!/bin/bash
# ncu --set full capture of the tcgen05 filter kernel (3 launches = one bench step after warm-up)
mkdir -p gpurun_out
timeout 900 ncu --set full --import-source on --clock-control none -k regex:knn_tc_filter --launch-skip 9 --launch-count 3 \
  -f -o gpurun_out/knn_v2 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-aggregation > gpurun_out/ncu_knn.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_knn.log; ls -la gpurun_out/*.ncu-rep
