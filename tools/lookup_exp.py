"""Lookup kernel experiment: pipeline depth x resident CTAs per SM at B=16, 512^2 (CUDA events)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import stitch_b200 as sb
from stitch_b200 import corr as C
from kernel_bench import timeit

B, n = 16, 4096
g = torch.Generator(device="cuda").manual_seed(0)
f1 = torch.randn(B, 256, 64, 64, device="cuda", generator=g)
f2 = torch.randn(B, 256, 64, 64, device="cuda", generator=g)
vol = C.corr(f1, f2)
maps = vol.view(B * n, 1, 64, 64)
coords = sb.lookup.coords_grid(B, 64, 64, device="cuda") + torch.randn(B, 2, 64, 64, device="cuda", generator=g) * 2
lib = sb._lib.load()
ref = None
for depth, per_sm, sbq in [(1, 3, 32), (1, 3, 16), (1, 2, 32), (2, 2, 32), (2, 2, 16)]:
    if True:
        lib.sb_tune(0, depth); lib.sb_tune(1, per_sm); lib.sb_tune(2, sbq)
        try:
            out = sb.encode_flow_token(maps, coords)
            torch.cuda.synchronize()
        except Exception as e:
            print(f"depth {depth} ctas/SM {per_sm}: {e}"); continue
        if ref is None:
            ref = out.clone()
        same = bool(torch.equal(out, ref))
        ms = timeit(lambda: sb.encode_flow_token(maps, coords), n=50)
        print(f"depth {depth} ctas/SM {per_sm} sbq {sbq}: {ms*1e3:7.1f} us  {B*n*732/ms/1e6:7.0f} GB/s alg  same={same}", flush=True)
lib.sb_tune(0, 0); lib.sb_tune(1, 0); lib.sb_tune(2, 0)
# L2 residency between iterations: 12 different coordinate sets around the same grid (as bench.py draws them)
cs = [sb.lookup.coords_grid(B, 64, 64, device="cuda") + torch.randn(B, 2, 64, 64, device="cuda", generator=g) * 2 for _ in range(12)]
def iters():
    for c in cs:
        sb.encode_flow_token(maps, c)
for eighths in (15, 2, 3, 4, 8):
    lib.sb_tune(4, eighths)
    ms = timeit(iters, n=10) / 12
    print(f"L2 keep {eighths}/8: {ms*1e3:7.1f} us per lookup (12 iterations, coords redrawn)", flush=True)
lib.sb_tune(4, 0)
for depth, per_sm in [(2, 2)]:
    lib.sb_tune(0, depth); lib.sb_tune(1, per_sm); lib.sb_tune(3, 1)
    ms = timeit(lambda: sb.encode_flow_token(maps, coords), n=50)
    print(f"FETCH ONLY depth {depth} ctas/SM {per_sm}: {ms*1e3:7.1f} us", flush=True)
lib.sb_tune(0, 0); lib.sb_tune(1, 0); lib.sb_tune(3, 0)
