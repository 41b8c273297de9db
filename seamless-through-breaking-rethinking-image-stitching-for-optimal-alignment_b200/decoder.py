"""N2 — convex 8x flow upsampling ("next" row 2 of SURVEY §8f).

Mirrors ``MemoryDecoder.upsample_flow(self, flow, mask)`` of the reference
(``core/FlowFormer/PerCostFormer3/decoder.py:214-225``).  The reference calls it after every
GRU iteration (``:331``) although evaluation only uses the last prediction; a caller that skips
the unused 11 calls saves 11 x 151 MB of mask reads per forward pass at 512^2, batch 16.
"""
from __future__ import annotations

import torch

from . import _lib

__all__ = ["upsample_flow", "memory_decoder_upsample_flow"]


def upsample_flow(flow, mask):
    """flow ``[N,2,H,W]``, mask ``[N,576,H,W]`` -> ``[N,2,8H,8W]`` (decoder.py:214-225)."""
    lib = _lib.load()
    fl = _lib.dev_f32(flow, "flow")
    mk = _lib.dev_f32(mask, "mask")
    if fl.dim() != 4 or fl.shape[1] != 2:
        raise ValueError(f"upsample_flow: flow must be [N,2,H,W], got {tuple(fl.shape)}")
    n, _, h, w = fl.shape
    if mk.shape != (n, 576, h, w):
        raise ValueError(f"upsample_flow: mask must be [{n},576,{h},{w}], got {tuple(mk.shape)}")
    out = torch.empty((n, 2, 8 * h, 8 * w), dtype=torch.float32, device=fl.device)
    _lib.check(lib.sb_upsample_flow(_lib.ptr(fl), _lib.ptr(mk), _lib.ptr(out), n, h, w, _lib.stream_ptr()),
               "sb_upsample_flow")
    return out


def memory_decoder_upsample_flow(self, flow, mask):
    """Drop-in body for ``MemoryDecoder.upsample_flow(self, flow, mask)``."""
    return upsample_flow(flow, mask)
