"""ncu target: GMA attention (fused two-pass softmax) at 16 x 4096 tokens, three calls."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import stitch_b200 as sb
B = 16
g = torch.Generator(device="cuda").manual_seed(0)
fm = torch.randn(B, 128, 64, 64, device="cuda", generator=g)
w_qk = torch.randn(256, 128, 1, 1, device="cuda", generator=g) * 0.02
for _ in range(3):
    a = sb.gma.attention(fm, w_qk)
torch.cuda.synchronize()
print("ok", float(a.sum()) / (B * 4096))
