"""Dev helper: the bench's PCA workload (2048 x 49152 -> 1024), a few calls -- ncu target."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from revisit_anything_b200 import engine
dev = torch.device("cuda")
S, Din, Dout = int(os.environ.get("PS", 2048)), 49152, 1024
g = torch.Generator(device=dev).manual_seed(12)
X = torch.randn(S, Din, generator=g, device=dev, dtype=torch.float64) / Din ** 0.5
W = torch.randn(Dout, Din, generator=g, device=dev) / Din ** 0.5
mu = torch.randn(Din, generator=g, device=dev, dtype=torch.float64) * 1e-3
ev = torch.rand(Dout, generator=g, device=dev) * 1e-4 + 1e-5
for _ in range(int(os.environ.get("CALLS", 3))):
    y = engine.pca_project(X, W, mu, ev, normalize_rows=True)
torch.cuda.synchronize()
print("ok", tuple(y.shape))
