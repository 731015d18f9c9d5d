"""Dev helper: the bench's fused aggregation + PCA workload (16 images x 128 SuperSegments, K=32 x 1536 -> 1024) -- ncu target."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from revisit_anything_b200 import engine
dev = torch.device("cuda")
B, N, D, K, S, Dout = 16, 1530, 1536, 32, 128, 1024
centers, tok, member, bits = bench._agg_workload(dev, B, N, D, K, S, seed=13)
g = torch.Generator(device=dev).manual_seed(14)
W = torch.randn(Dout, K * D, generator=g, device=dev) / (K * D) ** 0.5
mu = torch.randn(K * D, generator=g, device=dev, dtype=torch.float64) * 1e-3
ev = torch.rand(Dout, generator=g, device=dev) * 1e-4 + 1e-5
for _ in range(int(os.environ.get("CALLS", 3))):
    y = engine.aggregate_project_pca(tok, N, D, 0, centers, bits, [S] * B, None, W, mu, ev, normalize_rows=True)
torch.cuda.synchronize()
print("ok", tuple(y.shape))
