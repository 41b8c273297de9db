#!/usr/bin/env python
"""Times the REFERENCE'S OWN functions for the hot path on the host CPU.

Run by bench.py in a subprocess with CUDA_VISIBLE_DEVICES="" (mandatory: the reference moves tensors to CUDA
whenever one is visible — core/warp_utils.py:13-15, core/udis_utils/torch_homo_transform.py:46-53).  The
reference tree is the UNMODIFIED copy that `__graft_entry__.build()` places in baseline/_ref/ (git-ignored; it
travels to the GPU box with the gpurun snapshot).  Nothing from stitch_b200's kernels or the oracle runs here;
the only shared code is the synthetic-input generator (pipeline.make_pair_batch, plain torch on the CPU).

One step = one batch of pairs through the op list of SURVEY 8(d) config 2, every op being the reference function:
  2 x MemoryEncoder.corr (encoder.py:359-369) + the cost_maps view (encoder.py:260)
  2 x 3 F.avg_pool2d levels — C2 has no reference implementation; this is the RAFT form hinted at encoder.py:376
  24 x MemoryDecoder.encode_flow_token (decoder.py:242-260)
  FlowHomoAdpater.train_eval_foward (flowHomoAdpater.py:83-191) with stub networks: tensor_DLT, two
      torch_homo_transform.transformer calls, warp, overlap, compute_occlusion('wang'), threshold, multiply
Import shims: empty `skimage`, import-time-only fakes of `timm`, `.cuda()` -> identity.

Prints one JSON object: per-step seconds, pairs per step, thread count, torch parallel info, CPU model.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("STITCH_REF_COPY") or os.path.join(HERE, "_ref")


def install_shims(torch):
    sys.path[:0] = [REF, os.path.join(REF, "core")]
    for name in ("skimage", "skimage.io"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["skimage"].io = sys.modules["skimage.io"]

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    class _Dummy(torch.nn.Module):
        def __init__(self, *a, **k):
            super().__init__()

    timm = mod("timm", create_model=lambda *a, **k: None)
    mod("timm.data", IMAGENET_DEFAULT_MEAN=(0.485, 0.456, 0.406), IMAGENET_DEFAULT_STD=(0.229, 0.224, 0.225))
    layers = mod("timm.models.layers", Mlp=_Dummy, DropPath=_Dummy, to_2tuple=lambda x: (x, x),
                 trunc_normal_=lambda *a, **k: None, activations=types.SimpleNamespace())
    models = mod("timm.models", layers=layers)
    mod("timm.models.registry", register_model=lambda f: f)
    mod("timm.models.vision_transformer", Attention=_Dummy, Block=_Dummy, _cfg=lambda **k: {})
    timm.models = models
    torch.Tensor.cuda = lambda self, *a, **k: self          # flowHomoAdpater.py:92,101 guard on is_available(); belt and braces


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=16)
    ap.add_argument("--size", type=int, default=512)
    ap.add_argument("--iters", type=int, default=12)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--max-seconds", type=float, default=1e9, help="stop timing after this much timed work (>= 1 step)")
    ap.add_argument("--first-pair", type=int, default=0)
    args = ap.parse_args()

    if os.environ.get("CUDA_VISIBLE_DEVICES", None) != "":
        raise SystemExit("run_reference_cpu.py must run with CUDA_VISIBLE_DEVICES='' (the reference auto-moves to CUDA)")
    if not os.path.isdir(os.path.join(REF, "core")):
        raise SystemExit(f"{REF} is missing: run __graft_entry__.build() where /root/reference exists")
    import torch
    import torch.nn.functional as F
    assert not torch.cuda.is_available()
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    install_shims(torch)
    sys.path.insert(0, ROOT)
    from types import SimpleNamespace
    from core.FlowFormer.PerCostFormer3.encoder import MemoryEncoder
    from core.FlowFormer.PerCostFormer3.decoder import MemoryDecoder
    import core.flowHomoAdpater as ref_ad
    from stitch_b200.pipeline import make_pair_batch       # input generator only (CPU torch)

    class StubHomo(torch.nn.Module):
        def __init__(self, offsets):
            super().__init__()
            self.offsets = offsets

        def forward(self, a, b):
            return self.offsets.reshape(a.shape[0], -1), None

    class StubFlow(torch.nn.Module):
        def __init__(self, flows):
            super().__init__()
            self.flows, self.calls = flows, 0
            self.eval()

        def forward(self, a, b, out_dict=None):
            f = self.flows[self.calls % 2]
            self.calls += 1
            return [f.clone()]

    class Cfg:
        use_forward = False
        use_combine_h_flow = False
        use_fb_consistency_mask = True
        test_not_use_combine_h_flow = True
        only_homo = False

    pb = make_pair_batch(args.first_pair, args.pairs, size=args.size, iters=args.iters)
    enc_self = SimpleNamespace(cfg=SimpleNamespace(cost_heads_num=1))
    s8 = args.size // 8
    b = args.pairs
    adapter = ref_ad.FlowHomoAdpater(StubHomo(pb.h_motion), StubFlow([pb.flow_ij, pb.flow_ji]), Cfg()).eval()

    @torch.no_grad()
    def one_step():
        keep = []
        for (fa, fb), d in (((pb.fmap1, pb.fmap2), 0), ((pb.fmap2, pb.fmap1), 1)):
            vol = MemoryEncoder.corr(enc_self, fa, fb)
            cost_maps = vol.permute(0, 2, 3, 1, 4, 5).contiguous().view(b * s8 * s8, 1, s8, s8)      # encoder.py:260
            l1 = F.avg_pool2d(cost_maps, 2, stride=2)
            l2 = F.avg_pool2d(l1, 2, stride=2)
            l3 = F.avg_pool2d(l2, 2, stride=2)
            for it in range(args.iters):
                keep.append(MemoryDecoder.encode_flow_token(None, cost_maps, pb.coords[it, d]))
            keep.append(l3)
        adapter.flow_backbone.calls = 0
        out = adapter.train_eval_foward(pb.image1, pb.image2)
        return float(out["final_warp_output"].sum()) + float(keep[0].sum())

    for _ in range(args.warmup):
        one_step()
    per_step = []
    t_all = time.perf_counter()
    for _ in range(args.steps):
        t0 = time.perf_counter()
        one_step()
        per_step.append(time.perf_counter() - t0)
        if time.perf_counter() - t_all >= args.max_seconds:
            break
    cpu_model = "unknown"
    try:
        for line in open("/proc/cpuinfo"):
            if line.lower().startswith("model name"):
                cpu_model = line.split(":", 1)[1].strip()
                break
    except OSError:
        pass
    print(json.dumps({"step_seconds": per_step, "pairs_per_step": b, "threads": threads, "cpu_model": cpu_model,
                      "torch": torch.__version__, "parallel_info": torch.__config__.parallel_info().strip().splitlines()[:6],
                      "ref_files": "unmodified copy of the reference tree in baseline/_ref"}), flush=True)


if __name__ == "__main__":
    main()
