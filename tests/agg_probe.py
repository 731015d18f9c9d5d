"""Timing probe for the aggregate kernel (development aid, not a test): per-CTA cycle counters."""
import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from revisit_anything_b200 import _lib, engine, synth
lib = _lib.lib()
lib.segvlad_debug_aggregate_probe.argtypes = [C.c_void_p]
dev = torch.device("cuda")
B, N, D, K, S = 16, 1530, 1536, int(os.environ.get("PK", 64)), 128
g = torch.Generator(device=dev).manual_seed(11)
centers = synth.make_centers(K, D, 5).to(dev)
tok = torch.randn(B, D, N, generator=g, device=dev)
tok = tok / tok.norm(dim=1, keepdim=True) + 0.3 * (centers / centers.norm(dim=1, keepdim=True))[torch.randint(0, K, (B, N), generator=g, device=dev)].permute(0, 2, 1)
member = torch.rand(B * S, N, generator=g, device=dev) < float(os.environ.get("PRHO", 0.5))
bits = engine.pack_membership(member)
buf = torch.zeros(16 * 8192, dtype=torch.int64, device=dev)
for _ in range(2):
    engine.aggregate_batch(tok, N, D, 0, centers, bits, [S] * B, None)
lib.segvlad_debug_aggregate_probe(C.c_void_p(buf.data_ptr()))
engine.aggregate_batch(tok, N, D, 0, centers, bits, [S] * B, None)
torch.cuda.synchronize()
lib.segvlad_debug_aggregate_probe(None)
a = buf.cpu().numpy().reshape(-1, 16)
a = a[a[:, 0] > 0]
print("CTAs", len(a))
names = ["prod_total", "prod_poll", "prod_issue", "prod_slots", "cons_total", "cons_wait", "cons_epi", "cons_slots",
         "epi_flush+sums", "epi_bar1+norm(w0)", "epi_bar2", "rowloop(incl wait)"]
for i, n in enumerate(names):
    print(f"{n:12s} mean {a[:, i].mean():12.0f}  min {a[:, i].min():10d}  max {a[:, i].max():10d}")
print("cycles/slot (consumer total / slots):", (a[:, 4] / np.maximum(a[:, 7], 1)).mean())
