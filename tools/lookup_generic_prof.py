"""Full-size lookups (B=16, 512^2) through the TMA-box kernel and (SB_TUNE_LOOKUP_GENERIC) the LDG.128 window-staging
kernel, for an ncu comparison of the DRAM bytes the two access methods cost."""
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
cs = [sb.lookup.coords_grid(B, 64, 64, device="cuda") + torch.randn(B, 2, 64, 64, device="cuda", generator=g) * 2 for _ in range(3)]
for generic in (0, 1):
    lib.sb_tune(11, generic)
    for c in cs:
        out = sb.encode_flow_token(maps, c)
torch.cuda.synchronize()
print("ok")
