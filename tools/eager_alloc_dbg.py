"""Eager-mode HotPath.step: per-step host time and caching-allocator device allocations (is the step
host-bound on cudaMalloc while the allocator converges?)."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import stitch_b200 as sb
from stitch_b200.pipeline import HotPath, make_pair_batch
pb = make_pair_batch(0, 16, size=512, iters=12).map(lambda t: t.cuda())
hp = HotPath(size=512, iters=12, pyramid=True, overlap=True, eval_outputs=True)
out = None
for i in range(40):
    torch.cuda.synchronize()
    s0 = torch.cuda.memory_stats()
    t0 = time.perf_counter()
    out = hp.step(pb)
    th = time.perf_counter() - t0
    torch.cuda.synchronize()
    tt = time.perf_counter() - t0
    s1 = torch.cuda.memory_stats()
    print(f"step {i:2d}: enqueue {th*1e3:7.3f} ms, done {tt*1e3:7.3f} ms, cudaMalloc +{s1['num_device_alloc']-s0['num_device_alloc']}, cudaFree +{s1['num_device_free']-s0['num_device_free']}, "
          f"reserved {s1['reserved_bytes.all.current']/2**30:.2f} GiB, active {s1['active_bytes.all.current']/2**30:.2f} GiB", flush=True)
