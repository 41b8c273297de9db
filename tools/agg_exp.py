"""attn @ v kernel experiment: K steps per stage x L2 promotion (B=16, N=4096)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import stitch_b200 as sb
from kernel_bench import timeit
B, n = 16, 4096
g = torch.Generator(device="cuda").manual_seed(0)
attn = torch.rand(B, n, n, device="cuda", generator=g)
attn /= attn.sum(-1, keepdim=True)
vv = torch.randn(B, 128, n, device="cuda", generator=g)
fm = torch.randn(B, 128, n, device="cuda", generator=g)
gam = torch.tensor([0.5], device="cuda")
lib = sb._lib.load()
ref = None
for kps in (1, 2, 4, 0):
    for promo in (3,):
        lib.sb_tune(5, kps)
        out = sb.gma.attn_matmul_v(attn, vv, residual=fm, gamma=gam)
        torch.cuda.synchronize()
        if ref is None:
            ref = out.clone()
        ms = timeit(lambda: sb.gma.attn_matmul_v(attn, vv, residual=fm, gamma=gam), n=10)
        print(f"mblk {kps}: {ms*1e3:7.1f} us  {B*(n*n*4+3*n*128*4)/ms/1e6:6.0f} GB/s  maxdiff {float((out-ref).abs().max()):.2e}", flush=True)
