import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import stitch_b200 as sb
B = 16
g = torch.Generator(device="cuda").manual_seed(0)
f1 = torch.relu(torch.randn(B, 1024, 32, 32, device="cuda", generator=g))
f2 = torch.relu(torch.randn(B, 1024, 32, 32, device="cuda", generator=g))
for _ in range(3):
    out = sb.udis2_homography.CCL(f1, f2)
torch.cuda.synchronize()
print("ok")
