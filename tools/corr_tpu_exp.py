"""Experiment: target tiles per work unit of the cost-volume kernel (sb_tune key 13): 4 (default), 8, 16."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import stitch_b200 as sb
from stitch_b200 import corr as C
from kernel_bench import timeit
lib = sb._lib.load()
B = 16
g = torch.Generator(device="cuda").manual_seed(0)
f1 = torch.randn(B, 256, 64, 64, device="cuda", generator=g)
f2 = torch.randn(B, 256, 64, 64, device="cuda", generator=g)
t1, t2 = C.tokens_bf16(f1), C.tokens_bf16(f2)
ref, rlv = C.corr_from_tokens(t1, t2, 256, (64, 64), (64, 64), pyramid_levels=3)
ref = ref.clone(); rlv = [l.clone() for l in rlv]
for tpu in (4, 8, 16, 4):
    lib.sb_tune(13, tpu)
    v, lv = C.corr_from_tokens(t1, t2, 256, (64, 64), (64, 64), pyramid_levels=3)
    ok = torch.equal(v, ref) and all(torch.equal(a, b) for a, b in zip(lv, rlv))
    del v, lv
    ms3 = timeit(lambda: C.corr_from_tokens(t1, t2, 256, (64, 64), (64, 64), pyramid_levels=3))
    ms0 = timeit(lambda: C.corr_from_tokens(t1, t2, 256, (64, 64), (64, 64)))
    print(f"tiles per unit {tpu:2d}: bit-identical {ok}; with pyramid {ms3*1e3:.1f} us, without {ms0*1e3:.1f} us", flush=True)
lib.sb_tune(13, 0)
