"""Experiment: does a sub-batch's lookup get faster when its window lines are L2-resident?
12 lookups with re-drawn centres on the SAME n-pair volume, timed one by one, for a few launch configurations."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import stitch_b200 as sb
from stitch_b200 import corr as C
lib = sb._lib.load()
g = torch.Generator(device="cuda").manual_seed(0)
for npairs in (32, 16, 8):
    f1 = torch.randn(npairs, 256, 64, 64, device="cuda", generator=g)
    f2 = torch.randn(npairs, 256, 64, 64, device="cuda", generator=g)
    maps = C.corr(f1, f2).view(npairs * 4096, 1, 64, 64)
    coords = [sb.lookup.coords_grid(npairs, 64, 64, device="cuda") + torch.randn(npairs, 2, 64, 64, device="cuda", generator=g) * 2 for _ in range(12)]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for (depth, per_sm, sbq) in ((0, 0, 0), (1, 3, 8)):
        lib.sb_tune(0, depth); lib.sb_tune(1, per_sm); lib.sb_tune(2, sbq)
        # the 12 launches as one captured graph (no host time between them), replayed after an L2 flush / warm
        outs = [torch.empty((npairs, 64, 64, 81), device="cuda") for _ in range(12)]
        s_ = torch.cuda.Stream()
        s_.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s_):
            for i in range(12):
                sb.encode_flow_token(maps, coords[i], out=outs[i])
        torch.cuda.current_stream().wait_stream(s_)
        torch.cuda.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            for i in range(12):
                sb.encode_flow_token(maps, coords[i], out=outs[i])
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        res = []
        for cold in (True, False, False):
            if cold:
                flush.zero_()
            torch.cuda.synchronize()
            e0.record(); gr.replay(); e1.record()
            torch.cuda.synchronize()
            res.append(e0.elapsed_time(e1) * 1e3 / 12)
        print(f"pairs {npairs:2d} depth {depth} ctas/sm {per_sm} sbq {sbq}: per lookup in a 12-launch graph: cold L2 {res[0]:.2f} us, "
              f"warm {res[1]:.2f} / {res[2]:.2f} us  (= {res[2] * 16 / npairs:.1f} us per 16 pairs)", flush=True)
    lib.sb_tune(0, 0); lib.sb_tune(1, 0); lib.sb_tune(2, 0)
    del maps, f1, f2
