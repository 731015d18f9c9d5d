"""Dev helper: per-pass timeline of CTA 0 of aggregate_tc_kernel (TC_MARK probes in csrc/aggregate_tc.cu)."""
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
member = torch.rand(B * S, N, generator=g, device=dev) < 0.5
bits = engine.pack_membership(member)
buf = torch.zeros(4096 + 8 * 256, dtype=torch.int64, device=dev)
for _ in range(2):
    engine.aggregate_batch(tok, N, D, 0, centers, bits, [S] * B, None)
lib.segvlad_debug_aggregate_probe(C.c_void_p(buf.data_ptr()))
engine.aggregate_batch(tok, N, D, 0, centers, bits, [S] * B, None)
torch.cuda.synchronize()
lib.segvlad_debug_aggregate_probe(None)
t = buf.cpu().numpy()[4096:].reshape(256, 8)
t0 = t[0, 3] if t[0, 3] else t[0, 0]
names = ["mma:buf_free", "mma:stage_full", "mma:committed", "prod:issued", "epi:acc_full", "epi:released", "epi:write_beg", "epi:write_end"]
print("pass " + " ".join(f"{n:>15s}" for n in names) + "   (cycles since the first TMA issue; passes 0-11 norm sweep, 12-23 write sweep, ...)")
for i in range(int(os.environ.get("NPASS", 60))):
    print(f"{i:4d} " + " ".join(f"{int(t[i, j] - t0) if t[i, j] else 0:15d}" for j in range(8)))
