"""Launches each hot kernel twice at BASELINE config-2 size; run under ncu (-k regex ... --profile-from-start off):
the first repetition is a warm-up outside the profiled range (cudaProfilerStart / Stop around the second)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import stitch_b200 as sb
from stitch_b200 import corr as C

B, S = 16, 512
g = torch.Generator(device="cuda").manual_seed(0)
rnd = lambda *s: torch.randn(*s, device="cuda", generator=g)
f1, f2 = rnd(B, 256, 64, 64), rnd(B, 256, 64, 64)
x6 = torch.rand(B, 6, S, S, device="cuda", generator=g) * 255
img = torch.rand(B, 3, S, S, device="cuda", generator=g) * 255
flo = torch.nn.functional.interpolate(rnd(B, 2, 64, 64) * 2, size=(S, S), mode="bilinear", align_corners=True)
occ = (torch.rand(B, 1, S, S, device="cuda", generator=g) < 0.8).float()
src = sb.torch_DLT.corner_points(S, S, B, "cuda")
M = sb.torch_DLT.norm_matrix(S / 8, S / 8)
coords = sb.lookup.coords_grid(B, 64, 64, device="cuda") + rnd(B, 2, 64, 64) * 2
fm = rnd(B, 128, 64, 64)
w_qk, w_v, gam = rnd(256, 128, 1, 1) * 0.02, rnd(128, 128, 1, 1) * 0.09, torch.tensor([0.5], device="cuda")
um = rnd(B, 576, 64, 64)
cf1, cf2 = torch.relu(rnd(B, 1024, 32, 32)), torch.relu(rnd(B, 1024, 32, 32))
ys13, xs13 = torch.meshgrid(torch.linspace(-1, 1, 13, device="cuda"), torch.linspace(-1, 1, 13, device="cuda"), indexing="ij")
tps_src = torch.stack([xs13, ys13], -1).reshape(1, -1, 2).repeat(B, 1, 1)
tps_tgt = tps_src + 0.02 * rnd(B, 169, 2)
k_src = tps_src * 0.48 + 0.5
k_w, k_a = 0.01 * rnd(B, 169, 2), torch.tensor([[0.01, -0.02], [1.0, 0.01], [-0.01, 1.0]], device="cuda").repeat(B, 1, 1)
gw = lambda o, c: ((torch.rand(o, c, 6, 6, device="cuda", generator=g) * 2 - 1) / (c * 36) ** 0.5, (torch.rand(o, device="cuda", generator=g) * 2 - 1) / (c * 36) ** 0.5)
(pw1, pb1), (pw2, pb2), (pw3, pb3) = gw(16, 1), gw(32, 16), gw(64, 32)
pe_pack = sb.encoder.pack_patch_embed_weights(pw1, pw2, pw3)
for rep in range(2):
    if rep == 1:
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
    t1, t2 = C.tokens_bf16(f1), C.tokens_bf16(f2)
    vol, lv = C.corr_from_tokens(t1, t2, 256, (64, 64), (64, 64), pyramid_levels=3)
    vol0 = C.corr_from_tokens(t1, t2, 256, (64, 64), (64, 64))
    tok = sb.encode_flow_token(vol.view(B * 4096, 1, 64, 64), coords)
    pe = sb.encoder.patch_embed_proj(vol0.view(B * 4096, 1, 64, 64), pw1, pb1, pw2, pb2, pw3, pb3, pack=pe_pack)
    del pe
    H, th, thi = sb.torch_DLT.dlt_thetas(src / 8, (src + 5.0) / 8, left=sb.torch_DLT._inv3(M), right=M)
    oh = sb.torch_homo_transform.transformer(img, th, (S, S), append_ones=3)
    o = sb.compute_occlusion(flo, flo, "wang", occlusion_are_zeros=True, threshold=True)
    fw, ov = sb.warp(x6, flo, mul_mask=occ, return_overlap=True)
    mo = sb.preprocess_occlusion_mask(occ)
    attn = sb.gma.attention(fm, w_qk)
    agg = sb.gma.aggregate(attn, fm, w_v, gam)
    del attn
    up = sb.decoder.upsample_flow(coords - sb.lookup.coords_grid(B, 64, 64, device="cuda"), um)
    cc = sb.udis2_homography.CCL(cf1, cf2)
    tp = sb.torch_tps_transform.transformer(x6, tps_src, tps_tgt, (S, S))
    tk = sb.kornia_tps.warp_image_tps(x6, k_src, k_w, k_a)
    torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok")
