"""Cost volume: one CTA per tile vs CTA pairs (cta_group::2). Correctness (bit-compare) + timing."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import stitch_b200 as sb
from stitch_b200 import corr as C
from kernel_bench import timeit
lib = sb._lib.load()
g = torch.Generator(device="cuda").manual_seed(0)
def run(B, hw, lv, dtype=torch.float32):
    f1 = torch.randn(B, 256, *hw, device="cuda", generator=g); f2 = torch.randn(B, 256, *hw, device="cuda", generator=g)
    t1, t2 = C.tokens_bf16(f1), C.tokens_bf16(f2)
    res = {}
    for mode in (1, 2):
        lib.sb_tune(6, mode)
        try:
            out = C.corr_from_tokens(t1, t2, 256, hw, hw, pyramid_levels=lv, out_dtype=dtype)
            torch.cuda.synchronize()
        except Exception as e:
            print(f"mode {mode} B={B} hw={hw} lv={lv}: FAILED {str(e)[:120]}  dbg=0x{lib.sb_debug_word():08x}"); lib.sb_tune(6, 0); return False
        res[mode] = out
    lib.sb_tune(6, 0)
    a, b = res[1], res[2]
    if lv:
        same = torch.equal(a[0], b[0]) and all(torch.equal(x, y) for x, y in zip(a[1], b[1]))
    else:
        same = torch.equal(a, b)
    print(f"B={B} hw={hw} lv={lv} {dtype}: pair == single: {same}", flush=True)
    return same
ok = run(1, (16, 16), 0) and run(1, (16, 24), 0) and run(2, (64, 64), 0) and run(2, (64, 64), 3) and run(1, (9, 20), 0) and run(2, (64, 64), 0, torch.bfloat16)
print("correct:", ok)
if ok:
    B, n = 16, 4096
    f1 = torch.randn(B, 256, 64, 64, device="cuda", generator=g); f2 = torch.randn(B, 256, 64, 64, device="cuda", generator=g)
    t1, t2 = C.tokens_bf16(f1), C.tokens_bf16(f2)
    for mode in (1, 2):
        lib.sb_tune(6, mode)
        for lv, dt in ((0, torch.float32), (3, torch.float32), (0, torch.bfloat16)):
            ms = timeit(lambda: C.corr_from_tokens(t1, t2, 256, (64, 64), (64, 64), pyramid_levels=lv, out_dtype=dt))
            print(f"mode {mode} lv={lv} {dt}: {ms*1e3:7.1f} us  {B*2*n*n*256/ms/1e9:6.0f} TFLOP/s", flush=True)
    lib.sb_tune(6, 0)
