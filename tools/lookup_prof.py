"""Full-size lookups (B=16, 512^2), coords redrawn per call, for an ncu capture of corr_lookup_r4_kernel.
argv[1] = L2 keep eighths (sb_tune key 4)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import stitch_b200 as sb
from stitch_b200 import corr as C
B, n = 16, 4096
g = torch.Generator(device="cuda").manual_seed(0)
f1 = torch.randn(B, 256, 64, 64, device="cuda", generator=g)
f2 = torch.randn(B, 256, 64, 64, device="cuda", generator=g)
maps = C.corr(f1, f2).view(B * n, 1, 64, 64)
lib = sb._lib.load()
lib.sb_tune(4, int(sys.argv[1]) if len(sys.argv) > 1 else 0)
cs = [sb.lookup.coords_grid(B, 64, 64, device="cuda") + torch.randn(B, 2, 64, 64, device="cuda", generator=g) * 2 for _ in range(6)]
for c in cs:
    out = sb.encode_flow_token(maps, c)
torch.cuda.synchronize()
print("ok")
