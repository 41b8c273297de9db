import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests", "golden")]
import stitch_b200 as sb, stitch_oracle as so, cases
c = cases.ccl_small()
g = np.load(os.path.join(ROOT, "tests/golden/ccl_small.npz"))
flow = sb.udis2_homography.CCL(c["feature_1"].cuda(), c["feature_2"].cuda()).cpu().numpy()
print("small vs golden", np.abs(flow - g["flow"]).max())
gen = torch.Generator().manual_seed(82)
f1 = torch.relu(torch.randn(2, 1024, 32, 32, generator=gen))
f2 = torch.roll(f1, shifts=(2, -3), dims=(2, 3)) + 0.5 * torch.relu(torch.randn(2, 1024, 32, 32, generator=gen))
for scale in (10.0, 1.0):
    got = sb.udis2_homography.CCL(f1.cuda(), f2.cuda(), softmax_scale=scale).cpu().numpy()
    ref = so.ccl(f1.numpy(), f2.numpy(), softmax_scale=scale)
    print("full scale", scale, "max err", np.abs(got - ref).max(), "mean err", np.abs(got - ref).mean(), "ref range", ref.min(), ref.max())
