"""ncu target: three launches of the UDIS TPS warp kernel (16 x 512^2, 169 control points)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import stitch_b200 as sb
lib = sb._lib.load()
if len(sys.argv) > 1:
    lib.sb_tune(8, int(sys.argv[1]))
B, S = 16, 512
g = torch.Generator(device="cuda").manual_seed(0)
x6 = torch.rand(B, 6, S, S, device="cuda", generator=g) * 255
ys, xs = torch.meshgrid(torch.linspace(-1, 1, 13, device="cuda"), torch.linspace(-1, 1, 13, device="cuda"), indexing="ij")
sp = torch.stack([xs, ys], -1).reshape(1, -1, 2).repeat(B, 1, 1)
tg = sp + 0.02 * torch.randn(B, 169, 2, device="cuda", generator=g)
for _ in range(3):
    out = sb.torch_tps_transform.transformer(x6, sp, tg, (S, S))
torch.cuda.synchronize()
print("ok", float(out.mean()))
