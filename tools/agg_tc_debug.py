"""Dev helper: one small aggregation call through the tensor-core path, compared with the SIMT path."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from revisit_anything_b200 import engine

B, N, D, K, S = int(os.environ.get("B", 2)), int(os.environ.get("N", 150)), int(os.environ.get("D", 64)), int(os.environ.get("K", 8)), int(os.environ.get("S", 20))
g = torch.Generator(device="cuda").manual_seed(1)
centers = torch.randn(K, D, generator=g, device="cuda")
tok = torch.randn(B, N, D, generator=g, device="cuda")
member = torch.rand(B * S, N, generator=g, device="cuda") < 0.4
bits = engine.pack_membership(member)
os.environ["SEGVLAD_AGG_TC"] = "0"
ref = engine.aggregate_batch(tok, N, D, 1, centers, bits, [S] * B, None, out_dtype=torch.float64)
torch.cuda.synchronize()
os.environ["SEGVLAD_AGG_TC"] = "1"
out = engine.aggregate_batch(tok, N, D, 1, centers, bits, [S] * B, None, out_dtype=torch.float64)
torch.cuda.synchronize()
err = (out - ref).abs().max().item()
print("max abs diff tc vs simt:", err, "ref max", ref.abs().max().item())
