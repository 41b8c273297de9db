import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import stitch_b200 as sb
from stitch_b200 import corr as C
from kernel_bench import timeit
lib = sb._lib.load()
g = torch.Generator(device="cuda").manual_seed(0)
B = 16
f1 = torch.randn(B, 256, 64, 64, device="cuda", generator=g); f2 = torch.randn(B, 256, 64, 64, device="cuda", generator=g)
t1, t2 = C.tokens_bf16(f1), C.tokens_bf16(f2)
ref = C.corr_from_tokens(t1, t2, 256, (64, 64), (64, 64), pyramid_levels=3)
for pol in (3, 1, 5):
    lib.sb_tune(7, pol)
    out = C.corr_from_tokens(t1, t2, 256, (64, 64), (64, 64), pyramid_levels=3)
    same = torch.equal(out[0], ref[0]) and all(torch.equal(a, b) for a, b in zip(out[1], ref[1]))
    for lv in (0, 3):
        ms = timeit(lambda: C.corr_from_tokens(t1, t2, 256, (64, 64), (64, 64), pyramid_levels=lv))
        print(f"store policy {pol} lv={lv}: {ms*1e3:7.1f} us  same={same}", flush=True)
lib.sb_tune(7, 0)
