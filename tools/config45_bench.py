"""BASELINE.json configs 4 and 5 as timings (they are parity-test cases, not bench lines):
4: high-res 1024x1024 pairs (N = 16384 tokens, 1 GiB volume per pair and direction)
5: dense-flow backward warp + mask compositing sweep 256^2 .. 2048^2."""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import stitch_b200 as sb
from stitch_b200 import corr as C
from kernel_bench import timeit, report

g = torch.Generator(device="cuda").manual_seed(0)
rnd = lambda *s: torch.randn(*s, device="cuda", generator=g)
print("# config 4: 1024^2 pairs")
for B in (1, 2):
    f1, f2 = rnd(B, 256, 128, 128), rnd(B, 256, 128, 128)
    t1, t2 = C.tokens_bf16(f1), C.tokens_bf16(f2)
    n = 128 * 128
    ms = timeit(lambda: C.corr_from_tokens(t1, t2, 256, (128, 128), (128, 128)), n=5)
    report(f"corr_umma 1024^2 B={B}", ms, B * (n * n * 4 + 2 * n * 256 * 2))
    print(f"{'':34s} tensor: {B*2*n*n*256/ms/1e9:.0f} TFLOP/s")
    ms = timeit(lambda: C.corr_from_tokens(t1, t2, 256, (128, 128), (128, 128), pyramid_levels=3), n=3)
    report(f"  + fused 3-level pyramid B={B}", ms, B * (n * n * 4 * 1.328125 + 2 * n * 256 * 2))
    vol = C.corr_from_tokens(t1, t2, 256, (128, 128), (128, 128))
    coords = sb.lookup.coords_grid(B, 128, 128, device="cuda") + rnd(B, 2, 128, 128) * 2
    ms = timeit(lambda: sb.encode_flow_token(vol.view(B * n, 1, 128, 128), coords), n=10)
    report(f"corr_lookup r=4 1024^2 B={B}", ms, B * n * 732)
    del vol
print("# config 5: flow warp + test_out compositing sweep (B = 1)")
for S in (256, 512, 1024, 2048):
    x6 = torch.rand(1, 6, S, S, device="cuda", generator=g) * 255
    flo = rnd(1, 2, S, S) * 4
    occ = (torch.rand(1, 1, S, S, device="cuda", generator=g) < 0.8).float()
    px = S * S
    report(f"flow_warp C=6 {S}^2", timeit(lambda: sb.warp(x6, flo), n=20), px * 56)
    h1, h2 = x6, x6.flip(1)
    fw = sb.warp(x6, flo)
    report(f"composite_test_out {S}^2", timeit(lambda: sb.composite_test_out(h1, h2, fw, occ), n=20), px * 119)
