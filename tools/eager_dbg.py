import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import stitch_b200 as sb
from stitch_b200.pipeline import HotPath, make_pair_batch
from stitch_b200 import corr as C
pb = make_pair_batch(0, 16, size=512, iters=12).map(lambda t: t.cuda())
for overlap in (False, True):
    for ev in (False, True):
        hp = HotPath(size=512, iters=12, pyramid=True, overlap=overlap, eval_outputs=ev)
        for _ in range(3): hp.step(pb)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter(); e0.record()
        for _ in range(10): out = hp.step(pb)
        e1.record(); th = time.perf_counter() - t0
        torch.cuda.synchronize()
        print(f"overlap={overlap} eval_outputs={ev}: gpu {e0.elapsed_time(e1)/10:.3f} ms/step, host enqueue {th*100:.3f} ms/step", flush=True)
t1, t2 = C.tokens_bf16(pb.fmap1), C.tokens_bf16(pb.fmap2)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(20): C.corr_from_tokens(t1, t2, 256, (64, 64), (64, 64), pyramid_levels=3)
th = time.perf_counter() - t0
torch.cuda.synchronize()
print(f"corr_from_tokens host enqueue {th/20*1e3:.3f} ms/call")
vol, lv = C.corr_from_tokens(t1, t2, 256, (64, 64), (64, 64), pyramid_levels=3)
maps = vol.view(16 * 4096, 1, 64, 64)
torch.cuda.synchronize()
for name, fn in (("encode_flow_token", lambda: sb.encode_flow_token(maps, pb.coords[0])),
                 ("tokens_bf16", lambda: C.tokens_bf16(pb.fmap1)),
                 ("warp", lambda: sb.warp(torch.cat([pb.image1, pb.image1], 1), pb.flow_ij)),
                 ("empty 1GiB", lambda: torch.empty((16, 4096, 4096), device="cuda"))):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(20): r = fn()
    th = time.perf_counter() - t0
    torch.cuda.synchronize()
    tt = time.perf_counter() - t0
    print(f"{name}: host enqueue {th/20*1e6:.1f} us/call, with sync {tt/20*1e6:.1f} us/call")
print(torch.cuda.memory_stats()["num_alloc_retries"], torch.cuda.memory_reserved() / 2**30)
