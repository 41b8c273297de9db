"""TPS warps (W3 / W3k): time per launch with the two log() variants (SB_TUNE_TPS_LOG) and how far the
default (lg2.approx * ln2) moves the sample coordinates from the libdevice-logf variant and from the fp64
CPU oracle.  16 x 512^2, 13 x 13 control points (Homography/network.py:9-10)."""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import stitch_b200 as sb
import stitch_oracle as so
from kernel_bench import timeit, report

SB_TUNE_TPS_LOG = 8
lib = sb._lib.load()
B, S = 16, 512
g = torch.Generator(device="cuda").manual_seed(0)
rnd = lambda *s: torch.randn(*s, device="cuda", generator=g)
x6 = torch.rand(B, 6, S, S, device="cuda", generator=g) * 255
px = B * S * S
ys, xs = torch.meshgrid(torch.linspace(-1, 1, 13, device="cuda"), torch.linspace(-1, 1, 13, device="cuda"), indexing="ij")
sp = torch.stack([xs, ys], -1).reshape(1, -1, 2).repeat(B, 1, 1)
tg = sp + 0.02 * rnd(B, 169, 2)

xs_t, ys_t = sb.torch_homo_transform.linspace_table(S, x6.device), sb.torch_homo_transform.linspace_table(S, x6.device)
T0 = sb.torch_tps_transform.solve_system(sp, tg)
out0 = torch.empty_like(x6)
P = sb._lib.ptr
def tps_only():
    sb._lib.check(lib.sb_tps_warp(P(x6), P(T0), P(sp), P(xs_t), P(ys_t), P(out0), None, B, 6, S, S, S, S, 169, sb._lib.stream_ptr()), "tps")
for mode, name in ((1, "libdevice logf"), (0, "lg2.approx * ln2 (default)")):
    lib.sb_tune(SB_TUNE_TPS_LOG, mode)
    report(f"tps_warp kernel only: {name}", timeit(tps_only, n=5), px * 48)
res = {}
for mode, name in ((1, "libdevice logf"), (0, "lg2.approx * ln2 (default)")):
    lib.sb_tune(SB_TUNE_TPS_LOG, mode)
    ms = timeit(lambda: sb.torch_tps_transform.transformer(x6, sp, tg, (S, S)), n=5)
    report(f"tps_warp pn=169, {name}", ms, px * 48)
    print(f"{'':34s} {px*169/ms/1e6:.1f} G basis evaluations/s")
    out, idx = sb.torch_tps_transform.transformer(x6, sp, tg, (S, S), return_indices=True)
    res[mode] = (out.clone(), idx.clone())
mism = (res[0][1] != res[1][1]).any(dim=1)
print(f"UDIS TPS: taps that floor differently between the two log variants: {mism.float().mean().item():.2e}")
ok = ~mism[:, None].expand_as(res[0][0])
print(f"          max |out diff| elsewhere: {((res[0][0] - res[1][0]).abs() * ok).max().item():.3e} (images U(0,255) noise)")

# oracle at 2 x 256^2 (the CPU side is O(px * pn) python/numpy)
sp2, tg2 = sp[:2].cpu(), tg[:2].cpu()
yy, xx = torch.meshgrid(torch.linspace(0, 3.0, 256), torch.linspace(0, 4.0, 256), indexing="ij")
U = torch.stack([torch.sin(xx) + yy, torch.cos(yy), xx * 0.2, torch.ones_like(xx), torch.ones_like(xx), torch.ones_like(xx)], 0)[None].repeat(2, 1, 1, 1).contiguous()
T = sb.torch_tps_transform.solve_system(sp2.cuda(), tg2.cuda())
ref, ridx = so.tps_transformer(U.numpy(), sp2.numpy(), tg2.numpy(), (256, 256), return_indices=True, T=T.cpu().numpy())
for mode in (1, 0):
    lib.sb_tune(SB_TUNE_TPS_LOG, mode)
    out, idx = sb.torch_tps_transform.transformer(U.cuda(), sp2.cuda(), tg2.cuda(), (256, 256), return_indices=True)
    mm = (idx.cpu().numpy() != ridx).any(axis=1)
    okk = ~np.repeat(mm[:, None], 6, 1)
    d = np.abs(np.where(okk, out.cpu().numpy(), 0) - np.where(okk, ref, 0)).max()
    print(f"vs oracle, log mode {mode}: floor mismatches {mm.mean():.2e} (test bound 2e-3), max |diff| elsewhere {d:.2e} (bound 1e-3)")

# kornia-style warp: the grid it samples at, vs the fp64 oracle grid
src = torch.stack(torch.meshgrid(torch.linspace(0.02, 0.98, 13), torch.linspace(0.02, 0.98, 13), indexing="ij")[::-1], -1).reshape(1, -1, 2).repeat(B, 1, 1)
kw = 0.01 * torch.randn(B, 169, 2)
aw = torch.tensor([[0.01, -0.02], [1.0, 0.01], [-0.01, 1.0]]).repeat(B, 1, 1) + 0.01 * torch.randn(B, 3, 2)
rgrid = so.tps_kornia_grid(src[:1].numpy(), kw[:1].numpy(), aw[:1].numpy(), S, S)
for mode, name in ((1, "libdevice logf"), (0, "lg2.approx * ln2 (default)")):
    lib.sb_tune(SB_TUNE_TPS_LOG, mode)
    ms = timeit(lambda: sb.kornia_tps.warp_image_tps(x6, src.cuda(), kw.cuda(), aw.cuda()), n=5)
    report(f"tps_kornia K=169, {name}", ms, px * 48)
    out, grid = sb.kornia_tps.warp_image_tps(x6[:1], src[:1].cuda(), kw[:1].cuda(), aw[:1].cuda(), return_grid=True)
    print(f"{'':34s} max |grid - fp64 oracle| = {np.abs(grid.cpu().numpy() - rgrid).max():.2e} (test bound 1e-5)")
lib.sb_tune(SB_TUNE_TPS_LOG, 0)
