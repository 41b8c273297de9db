"""C1 / C2 — all-pairs cost volume and its avg-pool pyramid on tcgen05.

Mirrors ``MemoryEncoder.corr`` (reference
``core/FlowFormer/PerCostFormer3/encoder.py:359-369``): same argument meaning,
same output shape ``[B, heads, H1, W1, H2, W2]`` fp32 contiguous, so that
``cost_volume.permute(0,2,3,1,4,5).contiguous().view(B*H1*W1, heads, H2, W2)``
(``encoder.py:260``) stays a free view for ``heads == 1``.
The contraction runs with bf16 operands and fp32 accumulation (tolerance of the
parity contract: 1e-2 relative to the volume's scale).
"""
from __future__ import annotations

import sys

import torch

from . import _lib

__all__ = ["corr_bidirectional", "corr_bidirectional_from_tokens", "corr", "corr_pyramid", "memory_encoder_corr", "tokens_bf16", "corr_from_tokens"]


def _workspace(nbytes: int, device) -> torch.Tensor:
    # 256-byte aligned by the caching allocator (512-byte granularity)
    return torch.empty(max(nbytes, 256), dtype=torch.uint8, device=device)


def _split_heads(fmap: torch.Tensor, heads: int):
    b, dim, h, w = fmap.shape
    if dim % heads:
        raise ValueError(f"channel dim {dim} not divisible by heads={heads}")
    # 'b (heads d) h w' is already heads-major in memory: a free view
    return fmap.view(b * heads, dim // heads, h * w), b, dim // heads, h, w


def corr(fmap1: torch.Tensor, fmap2: torch.Tensor, heads: int = 1, pyramid_levels: int = 0):
    """corr[b,h,i,j] = sum_d fmap1[b,(h d),i] * fmap2[b,(h d),j]  (no 1/sqrt(d) scale).

    Returns the volume ``[B, heads, H1, W1, H2, W2]``; with ``pyramid_levels`` in
    1..3 also returns the list of pooled levels (see :func:`corr_pyramid`).
    """
    lib = _lib.load()
    f1 = _lib.dev_f32(fmap1, "fmap1")
    f2 = _lib.dev_f32(fmap2, "fmap2")
    if f1.dim() != 4 or f2.dim() != 4 or f1.shape[:2] != f2.shape[:2]:
        raise ValueError(f"corr: expected [B,C,H,W] feature maps with equal B,C; got {tuple(f1.shape)} {tuple(f2.shape)}")
    if not 0 <= pyramid_levels <= 3:
        raise ValueError("pyramid_levels must be in 0..3")
    v1, b, d, h1, w1 = _split_heads(f1, heads)
    v2, _, _, h2, w2 = _split_heads(f2, heads)
    bh = b * heads
    n1, n2 = h1 * w1, h2 * w2
    if n2 % 4 and bh * n1 * n2 > 0:
        # token count not a multiple of 4 (e.g. 65 x 67 maps): rows padded to the TMA's 16-byte stride
        # granularity; the result is a strided view with the reference's shape
        if pyramid_levels:
            raise ValueError("corr: the fused pyramid needs H2, W2 multiples of 8")
        n2p = (n2 + 3) // 4 * 4
        t1 = tokens_bf16(v1.reshape(bh, d, h1, w1))
        t2 = tokens_bf16(v2.reshape(bh, d, h2, w2))
        volp = torch.empty((bh, n1, n2p), dtype=torch.float32, device=f1.device)
        rc = lib.sb_corr_tokens_pitched(_lib.ptr(t1), _lib.ptr(t2), _lib.ptr(volp), n2p, None, None, None,
                                        bh, d, h1, w1, h2, w2, _lib.stream_ptr())
        _lib.check(rc, "sb_corr_tokens_pitched")
        return torch.as_strided(volp, (b, heads, h1, w1, h2, w2),
                                (heads * n1 * n2p, n1 * n2p, w1 * n2p, n2p, w2, 1))
    vol = torch.empty((bh, n1, n2), dtype=torch.float32, device=f1.device)
    lv = [None, None, None]
    for l in range(pyramid_levels):
        lv[l] = torch.empty((bh * n1, 1, h2 >> (l + 1), w2 >> (l + 1)), dtype=torch.float32, device=f1.device)
    if bh * n1 * n2 > 0:
        ws_bytes = lib.sb_corr_workspace_bytes(bh, d, n1, n2)
        ws = _workspace(ws_bytes, f1.device)
        rc = lib.sb_corr(_lib.ptr(v1), _lib.ptr(v2), _lib.ptr(vol), _lib.ptr(lv[0]), _lib.ptr(lv[1]),
                         _lib.ptr(lv[2]), _lib.ptr(ws), ws.numel(), bh, d, h1, w1, h2, w2, _lib.stream_ptr())
        _lib.check(rc, "sb_corr")
    out = vol.view(b, heads, h1, w1, h2, w2)
    if pyramid_levels:
        return out, [x for x in lv[:pyramid_levels]]
    return out


def corr_pyramid(fmap1: torch.Tensor, fmap2: torch.Tensor, num_levels: int = 4):
    """RAFT-style pyramid of per-query cost maps (C2).

    Level 0 is ``cost_maps`` = the volume viewed ``[B*H1*W1, 1, H2, W2]``; level l is
    ``F.avg_pool2d(level l-1, 2, stride=2)`` over the target axes — the form hinted
    at ``encoder.py:376`` and consumed by ``common.py:245-248``. The reference has
    no live implementation; the pooling is fused into the GEMM epilogue.
    """
    if not 1 <= num_levels <= 4:
        raise ValueError("num_levels must be in 1..4")
    res = corr(fmap1, fmap2, heads=1, pyramid_levels=num_levels - 1)
    vol, lv = (res, []) if num_levels == 1 else res
    b, _, h1, w1, h2, w2 = vol.shape
    return [vol.view(b * h1 * w1, 1, h2, w2)] + lv


def memory_encoder_corr(self, fmap1, fmap2):
    """Drop-in body for ``MemoryEncoder.corr(self, fmap1, fmap2)``."""
    return corr(fmap1, fmap2, heads=int(self.cfg.cost_heads_num))


def tokens_bf16(fmap: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
    """fp32 ``[B,C,H,W]`` -> bf16 token-major ``[B, H*W, Cpad]`` (the MMA operand
    layout). Lets a caller convert each image's features once and build both the
    forward and the backward volume from them.  ``out``: a dense bf16 ``[B, H*W, Cpad]`` buffer to fill."""
    lib = _lib.load()
    f = _lib.dev_f32(fmap, "fmap")
    b, c, h, w = f.shape
    cpad = (c + 63) // 64 * 64
    if out is not None:
        if tuple(out.shape) != (b, h * w, cpad) or out.dtype != torch.bfloat16 or not out.is_contiguous() or out.device != f.device:
            raise ValueError(f"tokens_bf16: out must be a dense bf16 [{b},{h * w},{cpad}] tensor on {f.device}")
    tok = out if out is not None else torch.empty((b, h * w, cpad), dtype=torch.bfloat16, device=f.device)
    _lib.check(lib.sb_feat_to_tokens_bf16(_lib.ptr(f), _lib.ptr(tok), b, c, h * w, _lib.stream_ptr()),
               "sb_feat_to_tokens_bf16")
    return tok


def corr_from_tokens(tok1: torch.Tensor, tok2: torch.Tensor, c: int, hw1, hw2, pyramid_levels: int = 0,
                     out_dtype=torch.float32):
    """Volume (+ pyramid) from two token-major bf16 maps. ``out_dtype=torch.bfloat16`` (opt-in, not the
    reference's dtype; no pyramid) stores the volume in bf16: half the HBM bytes."""
    lib = _lib.load()
    (h1, w1), (h2, w2) = hw1, hw2
    b = tok1.shape[0]
    if tok1.dtype != torch.bfloat16 or tok2.dtype != torch.bfloat16 or not tok1.is_cuda:
        raise RuntimeError("corr_from_tokens: expected CUDA bf16 token maps from tokens_bf16()")
    n1 = h1 * w1
    if out_dtype == torch.bfloat16:
        if pyramid_levels:
            raise ValueError("corr_from_tokens: the bf16 volume has no fused pyramid")
        vol16 = torch.empty((b, n1, h2 * w2), dtype=torch.bfloat16, device=tok1.device)
        rc = lib.sb_corr_tokens_bf16out(_lib.ptr(tok1), _lib.ptr(tok2), _lib.ptr(vol16), b, c, h1, w1, h2, w2,
                                        _lib.stream_ptr())
        _lib.check(rc, "sb_corr_tokens_bf16out")
        return vol16.view(b, 1, h1, w1, h2, w2)
    vol = torch.empty((b, n1, h2 * w2), dtype=torch.float32, device=tok1.device)
    lv = [None, None, None]
    for l in range(pyramid_levels):
        lv[l] = torch.empty((b * n1, 1, h2 >> (l + 1), w2 >> (l + 1)), dtype=torch.float32, device=tok1.device)
    rc = lib.sb_corr_tokens(_lib.ptr(tok1), _lib.ptr(tok2), _lib.ptr(vol), _lib.ptr(lv[0]), _lib.ptr(lv[1]),
                            _lib.ptr(lv[2]), b, c, h1, w1, h2, w2, _lib.stream_ptr())
    _lib.check(rc, "sb_corr_tokens")
    out = vol.view(b, 1, h1, w1, h2, w2)
    return (out, lv[:pyramid_levels]) if pyramid_levels else out


def corr_bidirectional(fmap1: torch.Tensor, fmap2: torch.Tensor, pyramid_levels: int = 0):
    """Forward and backward cost volume of a batch of pairs in ONE launch (heads = 1, equal map sizes).

    The reference builds them in two FlowFormer calls with swapped inputs (``flowHomoAdpater.py:158,178``); both need
    the same two feature maps.  Returns ``vol [2B,1,H,W,H,W]`` — ``vol[:B] == corr(fmap1, fmap2)``,
    ``vol[B:] == corr(fmap2, fmap1)`` bit for bit — and, with ``pyramid_levels``, the pooled levels ``[2B*H*W,1,h,w]``."""
    lib = _lib.load()
    f1 = _lib.dev_f32(fmap1, "fmap1")
    f2 = _lib.dev_f32(fmap2, "fmap2")
    if f1.dim() != 4 or f1.shape != f2.shape:
        raise ValueError(f"corr_bidirectional: expected two [B,C,H,W] feature maps of equal shape; got {tuple(f1.shape)} {tuple(f2.shape)}")
    if not 0 <= pyramid_levels <= 3:
        raise ValueError("pyramid_levels must be in 0..3")
    b, c, h, w = f1.shape
    n = h * w
    if n % 4:
        raise ValueError("corr_bidirectional: H*W must be a multiple of 4 (use corr() for ragged maps)")
    cpad = (c + 63) // 64 * 64
    tok = torch.empty((2 * b, n, cpad), dtype=torch.bfloat16, device=f1.device)
    tokens_bf16(f1, out=tok[:b])
    tokens_bf16(f2, out=tok[b:])
    return corr_bidirectional_from_tokens(tok, c, (h, w), pyramid_levels)


def corr_bidirectional_from_tokens(tok_both: torch.Tensor, c: int, hw, pyramid_levels: int = 0):
    """:func:`corr_bidirectional` from ``tok_both [2B, H*W, Cpad]`` (image-1 token maps, then image-2's)."""
    lib = _lib.load()
    h, w = hw
    n = h * w
    if tok_both.dtype != torch.bfloat16 or not tok_both.is_cuda or tok_both.dim() != 3 or tok_both.shape[0] % 2 or tok_both.shape[1] != n:
        raise RuntimeError("corr_bidirectional_from_tokens: expected a CUDA bf16 [2B, H*W, Cpad] token tensor")
    b = tok_both.shape[0] // 2
    vol = torch.empty((2 * b, n, n), dtype=torch.float32, device=tok_both.device)
    lv = [None, None, None]
    for l in range(pyramid_levels):
        lv[l] = torch.empty((2 * b * n, 1, h >> (l + 1), w >> (l + 1)), dtype=torch.float32, device=tok_both.device)
    if b * n > 0:
        _lib.check(lib.sb_corr_tokens_bidir(_lib.ptr(tok_both), _lib.ptr(vol), _lib.ptr(lv[0]), _lib.ptr(lv[1]), _lib.ptr(lv[2]),
                                            b, c, h, w, _lib.stream_ptr()), "sb_corr_tokens_bidir")
    out = vol.view(2 * b, 1, h, w, h, w)
    return (out, lv[:pyramid_levels]) if pyramid_levels else out


class _CallableModule(type(sys)):
    """``stitch_b200.corr(fmap1, fmap2, heads=1)`` — the package-level name SURVEY §8(b) asks for — and the
    module ``stitch_b200.corr`` (``corr.corr``, ``corr.tokens_bf16`` ...) are the same object."""

    def __call__(self, *args, **kwargs):
        return corr(*args, **kwargs)


sys.modules[__name__].__class__ = _CallableModule
