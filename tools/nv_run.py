"""NetVLAD leg alone (for ncu launch lists): python tools/nv_run.py"""
import json, os, sys
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench
peaks, _ = bench._peaks()
print(json.dumps(bench.netvlad_side_bench(torch.device("cuda", 0), peaks)))
