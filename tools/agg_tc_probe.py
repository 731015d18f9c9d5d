"""Dev helper: per-CTA cycle counters of aggregate_tc_kernel (see the TC_TIMED_WAIT probes in csrc/aggregate_tc.cu)."""
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
buf = torch.zeros(16 * 1024, dtype=torch.int64, device=dev)
for _ in range(2):
    engine.aggregate_batch(tok, N, D, 0, centers, bits, [S] * B, None)
lib.segvlad_debug_aggregate_probe(C.c_void_p(buf.data_ptr()))
engine.aggregate_batch(tok, N, D, 0, centers, bits, [S] * B, None)
torch.cuda.synchronize()
lib.segvlad_debug_aggregate_probe(None)
a = buf.cpu().numpy().reshape(-1, 16)
a = a[a[:, 0] > 0]
print("CTAs", len(a))
names = {0: "epi total", 1: "epi wait tfull (norm sweep)", 3: "epi bar.sync", 4: "epi wait tfull (write sweep)",
         5: "epi ld+stage+store (write sweep)", 6: "mma wait tempty", 7: "mma wait full", 8: "mma total",
         9: "producer wait empty", 10: "builder wait empty", 11: "builder total", 12: "items", 13: "epi sibling exchange wait"}
for i, n in names.items():
    print(f"{n:36s} mean {a[:, i].mean():12.0f}  min {a[:, i].min():10d}  max {a[:, i].max():10d}")
