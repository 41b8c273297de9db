"""C3p: the four pyramid-level lookups one after the other vs on four streams (STITCH_B200_PYRAMID_LOOKUP_STREAMS=1 / 4)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import stitch_b200 as sb
from kernel_bench import timeit
g = torch.Generator(device="cuda").manual_seed(0)
B = 16
f1 = torch.randn(B, 256, 64, 64, device="cuda", generator=g); f2 = torch.randn(B, 256, 64, 64, device="cuda", generator=g)
coords = sb.lookup.coords_grid(B, 64, 64, device="cuda") + torch.randn(B, 2, 64, 64, device="cuda", generator=g) * 2
pyr = sb.corr_pyramid(f1, f2, 4)
out = sb.encode_flow_token_pyramid(pyr, coords); torch.cuda.synchronize()
# graph-timed (per-call event timing of four small launches is host-bound)
st = torch.cuda.Stream()
with torch.cuda.stream(st):
    for _ in range(3): sb.encode_flow_token_pyramid(pyr, coords)
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr, stream=st):
        for _ in range(20): keep = sb.encode_flow_token_pyramid(pyr, coords)
    gr.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); gr.replay(); e1.record(); torch.cuda.synchronize()
print("  graph replay of 20 calls: %.1f us per call" % (e0.elapsed_time(e1) / 20 * 1e3))
print("streams", os.environ.get("STITCH_B200_PYRAMID_LOOKUP_STREAMS", "4"), "pyramid lookup %.1f us" % (timeit(lambda: sb.encode_flow_token_pyramid(pyr, coords)) * 1e3), "checksum %.6e" % out.double().sum().item())
