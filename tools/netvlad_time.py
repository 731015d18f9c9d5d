"""Dev helper: timing of the NetVLAD anti-burst aggregation on BASELINE config 5's shape (128 centres x 768-D, 23 x 23 tokens)."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from revisit_anything_b200 import engine
dev = torch.device("cuda")
B, D, H, K = int(os.environ.get("PB", 256)), 768, 23, 128
g = torch.Generator(device=dev).manual_seed(5)
x = torch.randn(B, D, H * H, generator=g, device=dev)
cent = torch.rand(K, D, generator=g, device=dev)
W = 12.0 * cent / cent.norm(dim=1, keepdim=True)
for _ in range(2):
    y = engine.netvlad_antiburst(x, cent, W)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    y = engine.netvlad_antiburst(x, cent, W)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
N = H * H
flops = B * (2.0 * N * N * D + 2 * 2.0 * N * K * D)
print(f"netvlad antiburst: B={B} images, {ms:.3f} ms/batch, {B / ms * 1e3:.0f} images/s, {flops / ms / 1e9:.1f} TFLOP/s (self-similarity + soft-assign + aggregation)")
