"""Cost volume: A operand from shared memory (default) vs copied to tensor memory once per unit
(SB_TUNE_CORR_A_TMEM = 1): time and bit-identity, 16 x 4096^2, C = 256 and C = 128."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import stitch_b200 as sb
from stitch_b200 import corr as C
from kernel_bench import timeit, report
lib = sb._lib.load()
B, n = 16, 4096
g = torch.Generator(device="cuda").manual_seed(0)
for ch in (256, 128):
    f1, f2 = torch.randn(B, ch, 64, 64, device="cuda", generator=g), torch.randn(B, ch, 64, 64, device="cuda", generator=g)
    t1, t2 = C.tokens_bf16(f1), C.tokens_bf16(f2)
    ref = {}
    for mode in (2, 1, 2, 1):        # 2 = shared-memory A operand, 1 = A copied to tensor memory (default)
        lib.sb_tune(10, mode)
        for lv in (0, 3):
            out = C.corr_from_tokens(t1, t2, ch, (64, 64), (64, 64), pyramid_levels=lv)
            vol = out[0] if lv else out
            torch.cuda.synchronize()
            same = ""
            if mode == 0 and lv not in ref:
                ref[lv] = vol.clone()
            elif mode == 1:
                same = f"   identical: {bool(torch.equal(vol, ref[lv]))}" + ("" if torch.equal(vol, ref[lv]) else f" max diff {(vol - ref[lv]).abs().max().item():.3e}")
            del out, vol
            ms = timeit(lambda: C.corr_from_tokens(t1, t2, ch, (64, 64), (64, 64), pyramid_levels=lv), n=10)
            report(f"corr C={ch} levels={lv} a_tmem={mode}", ms, B * (n * n * 4 * (1.328125 if lv else 1) + 2 * n * ch * 2))
            if same: print(same)
lib.sb_tune(10, 0)
