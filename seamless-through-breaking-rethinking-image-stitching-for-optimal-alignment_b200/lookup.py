"""C3 / C3p — per-iteration bilinear correlation lookup.

Mirrors ``MemoryDecoder.encode_flow_token`` (reference
``core/FlowFormer/PerCostFormer3/decoder.py:242-260``) and ``bilinear_sampler`` /
``coords_grid`` (``core/utils/utils.py:62-76,97-100``).
"""
from __future__ import annotations

import torch

from . import _lib

__all__ = ["encode_flow_token", "encode_flow_token_pyramid", "memory_decoder_encode_flow_token",
           "bilinear_sampler", "coords_grid"]


def coords_grid(batch, ht, wd, device=None):
    """(x, y) integer pixel grid ``[B, 2, H, W]`` fp32 (utils.py:97-100)."""
    ys, xs = torch.meshgrid(torch.arange(ht, device=device), torch.arange(wd, device=device), indexing="ij")
    return torch.stack([xs, ys], dim=0).float()[None].repeat(batch, 1, 1, 1)


def bilinear_sampler(img, coords, mode="bilinear", mask=False):
    """grid_sample wrapper in pixel coordinates (utils.py:62-76).

    img ``[N,C,H,W]``, coords ``[N,Ho,Wo,2]`` (x,y) -> ``[N,C,Ho,Wo]``; zeros
    padding, align_corners=True. ``mask=True`` also returns the in-range mask
    computed exactly like the reference (strict inequalities on the normalised grid).
    """
    if mode != "bilinear":
        raise NotImplementedError("bilinear_sampler: only mode='bilinear' exists in the reference")
    lib = _lib.load()
    im = _lib.dev_f32(img, "img")
    co = _lib.dev_f32(coords, "coords")
    n, c, h, w = im.shape
    if co.dim() != 4 or co.shape[0] != n or co.shape[-1] != 2:
        raise ValueError(f"bilinear_sampler: coords {tuple(co.shape)} do not match img {tuple(im.shape)}")
    ho, wo = co.shape[1], co.shape[2]
    out = torch.empty((n, c, ho, wo), dtype=torch.float32, device=im.device)
    _lib.check(lib.sb_bilinear_sampler(_lib.ptr(im), _lib.ptr(co), _lib.ptr(out), n, c, h, w, ho, wo,
                                       _lib.stream_ptr()), "sb_bilinear_sampler")
    if mask:
        xg = 2 * co[..., 0:1] / (w - 1) - 1
        yg = 2 * co[..., 1:2] / (h - 1) - 1
        m = (xg > -1) & (yg > -1) & (xg < 1) & (yg < 1)
        return out, m.float()
    return out


def _lookup_into(cost_maps, coords, out, r, scale, stride, offset):
    lib = _lib.load()
    nq, heads, h2, w2 = cost_maps.shape
    b, _, h1, w1 = coords.shape
    rc = lib.sb_corr_lookup(_lib.ptr(cost_maps), _lib.ptr(coords), _lib.ptr(out), b, h1, w1, h2, w2, r,
                            float(scale), stride, offset, _lib.stream_ptr())
    _lib.check(rc, "sb_corr_lookup")


def encode_flow_token(cost_maps, coords, r=4, out=None):
    """cost_maps ``[B*H1*W1, heads, H2, W2]``, coords ``[B,2,H1,W1]`` (x,y) ->
    ``[B, heads*(2r+1)^2, H1, W1]`` with the reference's memory order
    ``[B,H1,W1,heads*(2r+1)^2]``; channel ``k = i*(2r+1)+j`` samples
    ``(cx + i - r, cy + j - r)`` (decoder.py:250-256, RAFT's meshgrid(dy,dx) quirk).

    ``out`` (extension, heads == 1): a dense ``[B,H1,W1,(2r+1)^2]`` fp32 buffer to write into — e.g. a batch
    slice of a larger result when the decoder loop runs per sub-batch of pairs."""
    cm = _lib.dev_f32(cost_maps, "cost_maps")
    co = _lib.dev_f32(coords, "coords")
    if cm.dim() != 4 or co.dim() != 4 or co.shape[1] != 2:
        raise ValueError(f"encode_flow_token: bad shapes {tuple(cm.shape)} {tuple(co.shape)}")
    b, _, h1, w1 = co.shape
    nq, heads, h2, w2 = cm.shape
    if nq != b * h1 * w1:
        raise ValueError(f"encode_flow_token: {nq} cost maps for {b}x{h1}x{w1} queries")
    side = 2 * r + 1
    if heads == 1:
        if out is None:
            out = torch.empty((b, h1, w1, side * side), dtype=torch.float32, device=cm.device)
        elif (tuple(out.shape) != (b, h1, w1, side * side) or out.dtype != torch.float32 or not out.is_contiguous()
              or out.device != cm.device):
            raise ValueError(f"encode_flow_token: out must be a dense fp32 [{b},{h1},{w1},{side * side}] tensor on {cm.device}")
        if nq:
            _lookup_into(cm, co, out, r, 1.0, side * side, 0)
        return out.permute(0, 3, 1, 2)
    if out is not None:
        raise ValueError("encode_flow_token: out= is only supported for heads == 1")
    # multi-head maps (not used by the shipped config, cost_heads_num=1): generic sampler
    d = torch.linspace(-r, r, side, device=cm.device)
    delta = torch.stack(torch.meshgrid(d, d, indexing="ij"), dim=-1).view(1, side, side, 2)
    pts = co.permute(0, 2, 3, 1).reshape(nq, 1, 1, 2) + delta
    samp = bilinear_sampler(cm, pts)
    return samp.view(b, h1, w1, -1).permute(0, 3, 1, 2)


def encode_flow_token_pyramid(cost_pyramid, coords, r=4):
    """C3p: lookup on every pyramid level with the centre divided by ``2**l`` and
    the same integer window (``coords = centroid / 2**i + delta``, reference dead
    code ``core/FlowFormer/common.py:245-248``). Returns ``[B, L*(2r+1)^2, H1, W1]``
    (memory ``[B,H1,W1,L*(2r+1)^2]``), level-major channels as RAFT concatenates."""
    co = _lib.dev_f32(coords, "coords")
    b, _, h1, w1 = co.shape
    side = 2 * r + 1
    nl = len(cost_pyramid)
    out = torch.empty((b, h1, w1, nl * side * side), dtype=torch.float32, device=co.device)
    levels = []
    for l, cm in enumerate(cost_pyramid):
        cm = _lib.dev_f32(cm, f"cost_pyramid[{l}]")
        if cm.shape[0] != b * h1 * w1 or cm.shape[1] != 1:
            raise ValueError(f"encode_flow_token_pyramid: level {l} has shape {tuple(cm.shape)}")
        levels.append(cm)
    if not out.numel():
        return out.permute(0, 3, 1, 2)
    # The levels are independent launches that write disjoint channel slices of `out`: with
    # STITCH_B200_PYRAMID_LOOKUP_STREAMS = n > 1 level l runs on stream l mod n (0 = the caller's; the others are forked
    # from it and joined before returning), so that the ramp and the tail of a launch are filled by another level's
    # CTAs.  Measured at B = 16 (tools/pyramid_lookup_exp.py): one stream 123 us eager / 115 us inside a captured graph,
    # two 111-116 / 103, four 162 / 104 (eagerly the extra event records cost more than the overlap gives) -> default 2.
    cur = torch.cuda.current_stream(co.device)
    aux = _aux_streams(co.device, nl - 1, cur.priority) if _PYR_STREAMS > 1 else []
    for st in aux:
        st.wait_stream(cur)                       # fork once: everything enqueued so far precedes every level
    for l, cm in enumerate(levels):
        k = l % (len(aux) + 1)
        st = aux[k - 1] if k > 0 else None
        if st is None:
            _lookup_into(cm, co, out, r, 1.0 / (1 << l), nl * side * side, l * side * side)
        else:
            with torch.cuda.stream(st):
                _lookup_into(cm, co, out, r, 1.0 / (1 << l), nl * side * side, l * side * side)
    for st in aux:
        cur.wait_stream(st)
    return out.permute(0, 3, 1, 2)


import os as _os

_PYR_STREAMS = int(_os.environ.get("STITCH_B200_PYRAMID_LOOKUP_STREAMS", "2"))
_AUX = {}


def _aux_streams(device, n, priority):
    key = (torch.device(device).index if torch.device(device).index is not None else torch.cuda.current_device(), priority)
    lst = _AUX.setdefault(key, [])
    while len(lst) < min(n, _PYR_STREAMS - 1):
        lst.append(torch.cuda.Stream(device, priority=priority))
    return lst[:min(n, _PYR_STREAMS - 1)]


def memory_decoder_encode_flow_token(self, cost_maps, coords, r=4):
    """Drop-in body for ``MemoryDecoder.encode_flow_token(self, cost_maps, coords, r=4)``."""
    return encode_flow_token(cost_maps, coords, r)
