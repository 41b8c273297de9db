"""Where the captured step's time goes: graph replays of HotPath with 12 / 6 / 0 lookup iterations, one or two lookup
chains, with and without the warp stage beside the cost stage (batch 16, 512^2)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import stitch_b200 as sb
from stitch_b200.pipeline import HotPath, make_pair_batch

def time_cfg(iters, streams, warp=True, reps=20):
    pb = make_pair_batch(0, 16, size=512, iters=max(iters, 1)).map(lambda t: t.cuda())
    hp = HotPath(size=512, iters=iters, pyramid=True)
    hp.lookup_streams = streams
    if not warp:
        hp._warp_stage = lambda pb_: {}
    hp.capture(pb)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); hp.replay(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2] * 1e3

for warp in (True, False):
    for streams in (2, 1):
        row = [time_cfg(it, streams, warp) for it in (12, 6, 0)]
        print("warp stage %-5s lookup chains %d: 12 iters %.0f us, 6 iters %.0f us, 0 iters %.0f us -> %.1f us per iteration (both directions)" % (
            warp, streams, row[0], row[1], row[2], (row[0] - row[2]) / 12), flush=True)
