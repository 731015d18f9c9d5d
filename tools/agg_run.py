"""Dev helper: the bench's aggregation workload (16 images x 128 SuperSegments, K=64, D=1536), a few calls -- ncu target."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from revisit_anything_b200 import engine, synth

dev = torch.device("cuda")
B, N, D, K, S = 16, 1530, 1536, int(os.environ.get("PK", 64)), 128
g = torch.Generator(device=dev).manual_seed(11)
centers = synth.make_centers(K, D, 5).to(dev)
tok = torch.randn(B, D, N, generator=g, device=dev)
tok = tok / tok.norm(dim=1, keepdim=True) + 0.3 * (centers / centers.norm(dim=1, keepdim=True))[
    torch.randint(0, K, (B, N), generator=g, device=dev)].permute(0, 2, 1)
member = torch.rand(B * S, N, generator=g, device=dev) < float(os.environ.get("PRHO", 0.5))
bits = engine.pack_membership(member)
for _ in range(int(os.environ.get("CALLS", 3))):
    out = engine.aggregate_batch(tok, N, D, 0, centers, bits, [S] * B, None, out_dtype=torch.float64)
torch.cuda.synchronize()
print("ok", tuple(out.shape))
