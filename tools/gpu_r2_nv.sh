#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_netvlad.py -m gpu -x -q 2>&1 | tail -25
timeout 300 python - <<'PY'
import sys, json, torch
sys.path.insert(0, ".")
import bench
peaks,_ = bench._peaks()
print(json.dumps(bench.netvlad_side_bench(torch.device("cuda",0), peaks)))
import os
os.environ["SEGVLAD_NETVLAD_TC"]="0"
print(json.dumps(bench.netvlad_side_bench(torch.device("cuda",0), peaks)))
PY
