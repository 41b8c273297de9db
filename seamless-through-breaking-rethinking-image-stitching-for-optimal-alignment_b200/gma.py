"""N1 — GMA attention and aggregation ("next" row 1 of SURVEY §8f).

Mirrors ``Attention`` / ``Aggregate`` of the reference's
``core/FlowFormer/PerCostFormer3/gma.py`` (:35-76, :79-115; built at ``decoder.py:197`` and
``gru.py:316``, called at ``decoder.py:283`` and, every GRU iteration, ``gru.py:324``).

* ``attention``: ``softmax((scale*q) . k^T)`` — the contraction is the cost volume's (K = 128) and
  runs on the tcgen05 correlation kernel; a row-softmax kernel normalises in place.
* ``aggregate``: ``fmap + gamma * (attn @ v)`` — a TF32 tcgen05 GEMM that consumes the fp32
  attention matrix directly (TMA -> swizzled smem -> ``tcgen05.mma.kind::tf32``), with the
  transposed store and the residual fused into its epilogue.
The 1x1 convolutions (``to_qk``, ``to_v``, ``project``) are plain library GEMMs (``F.conv2d``).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F
from torch import nn

from . import _lib
from . import corr as corr_mod

__all__ = ["attention", "aggregate", "project_qk", "attention_forward", "aggregate_forward", "Attention", "Aggregate", "softmax_rows_", "attn_matmul_v"]


def softmax_rows_(x: torch.Tensor, to_tf32: bool = True) -> torch.Tensor:
    """In-place softmax over the last dim of a dense fp32 CUDA tensor ``[..., n]``."""
    lib = _lib.load()
    if not (x.is_cuda and x.dtype == torch.float32 and x.is_contiguous()):
        raise ValueError("softmax_rows_: expected a contiguous fp32 CUDA tensor")
    n = x.shape[-1]
    rows = x.numel() // max(n, 1)
    _lib.check(lib.sb_softmax_rows(_lib.ptr(x), rows, n, n, 1 if to_tf32 else 0, _lib.stream_ptr()), "sb_softmax_rows")
    return x


def project_qk(fmap, to_qk_weight, heads: int = 1, scale: float | None = None):
    """The 1x1 ``to_qk`` convolution and head split (gma.py:57-60): ``scale*q, k`` as ``[b*heads, d, h, w]``."""
    fm = _lib.dev_f32(fmap, "fmap")
    b, c, h, w = fm.shape
    qk = F.conv2d(fm, to_qk_weight)
    q, k = qk.chunk(2, dim=1)
    d = q.shape[1] // heads
    scale = d ** -0.5 if scale is None else scale
    return (scale * q).reshape(b * heads, d, h, w), k.reshape(b * heads, d, h, w)


def attention(fmap, to_qk_weight, heads: int = 1, scale: float | None = None, dtype=torch.float32):
    """``Attention.forward`` (gma.py:54-76): fmap ``[b,c,h,w]`` -> attn ``[b, heads, h*w, h*w]``.

    ``dtype=torch.float32`` (default) is the reference's return type. ``dtype=torch.bfloat16`` is an
    opt-in: the probabilities are stored in bf16, which halves what every ``aggregate`` call (one per
    GRU iteration) has to read from HBM."""
    b, c, h, w = fmap.shape
    q, k = project_qk(fmap, to_qk_weight, heads, scale)
    n = h * w
    if dtype == torch.bfloat16:
        lib = _lib.load()
        sim = corr_mod.corr(q, k).view(b * heads, n, n)
        out = torch.empty((b * heads, n, n), dtype=torch.bfloat16, device=sim.device)
        _lib.check(lib.sb_softmax_rows_bf16(_lib.ptr(sim), _lib.ptr(out), b * heads * n, n, n, _lib.stream_ptr()),
                   "sb_softmax_rows_bf16")
        return out.view(b, heads, n, n)
    if dtype != torch.float32:
        raise ValueError("attention: dtype must be torch.float32 or torch.bfloat16")
    if n % 4 == 0:
        # two passes over the same bf16 x bf16 -> fp32 contraction (row statistics, then normalised
        # probabilities): the fp32 logits never travel through HBM
        lib = _lib.load()
        tq, tk = corr_mod.tokens_bf16(q), corr_mod.tokens_bf16(k)
        attn = torch.empty((b * heads, n, n), dtype=torch.float32, device=tq.device)
        stats = torch.empty((b * heads, n, 2), dtype=torch.float32, device=tq.device)
        _lib.check(lib.sb_attn_softmax_tokens(_lib.ptr(tq), _lib.ptr(tk), _lib.ptr(attn), _lib.ptr(stats), b * heads,
                                              q.shape[1], n, n, _lib.stream_ptr()), "sb_attn_softmax_tokens")
        return attn.view(b, heads, n, n)
    # token counts that are not a multiple of 4 (e.g. 65 x 67 maps): the contraction still runs on the tensor
    # cores (corr() pads the row pitch for the TMA and returns a strided view); the row softmax of that ragged,
    # pitched matrix is library code (torch), like the 1x1 convolutions
    sim = corr_mod.corr(q, k).reshape(b * heads, n, n)       # bf16 x bf16 -> fp32 on the tensor cores
    return torch.softmax(sim, dim=-1).view(b, heads, n, n)


def attn_matmul_v(attn, v, residual=None, gamma=None):
    """out[bh, n, i] = sum_j attn[bh, i, j] * v[bh, n, j]  (+ residual + gamma scaling fused).

    attn ``[BH, Nq, Nk]``, v ``[BH, d, Nk]`` (the conv layout), residual ``[BH, d, Nq]`` or None,
    gamma: 1-element CUDA tensor or None -> ``[BH, d, Nq]``."""
    lib = _lib.load()
    if attn.dtype == torch.bfloat16:
        if not attn.is_cuda:
            raise RuntimeError("attn_matmul_v: stitch_b200 runs on B200 GPUs only (no CPU fallback)")
        a = attn.contiguous()
        vv = v.to(torch.bfloat16).contiguous()
        bh, nq, nk = a.shape
        d = vv.shape[1]
        if vv.shape != (bh, d, nk):
            raise ValueError(f"attn_matmul_v: v {tuple(vv.shape)} does not match attn {tuple(a.shape)}")
        res = _lib.dev_f32(residual, "residual") if residual is not None else None
        gm = _lib.dev_f32(gamma, "gamma") if gamma is not None else None
        out = torch.empty((bh, d, nq), dtype=torch.float32, device=a.device)
        _lib.check(lib.sb_attn_aggregate_bf16(_lib.ptr(a), _lib.ptr(vv), _lib.ptr(res), _lib.ptr(gm), _lib.ptr(out),
                                              bh, nq, nk, d, _lib.stream_ptr()), "sb_attn_aggregate_bf16")
        return out
    a = _lib.dev_f32(attn, "attn")
    vv = _lib.dev_f32(v, "v")
    bh, nq, nk = a.shape
    d = vv.shape[1]
    if vv.shape != (bh, d, nk):
        raise ValueError(f"attn_matmul_v: v {tuple(vv.shape)} does not match attn {tuple(a.shape)}")
    res = _lib.dev_f32(residual, "residual") if residual is not None else None
    gm = _lib.dev_f32(gamma, "gamma") if gamma is not None else None
    if nk % 4:
        # ragged key count (the TMA needs 16-byte row strides): library GEMM on the GPU, same formula
        out = torch.bmm(vv, a.transpose(1, 2))
        if gm is not None:
            out = gm * out
        return out if res is None else res + out
    out = torch.empty((bh, d, nq), dtype=torch.float32, device=a.device)
    _lib.check(lib.sb_attn_aggregate(_lib.ptr(a), _lib.ptr(vv), _lib.ptr(res), _lib.ptr(gm), _lib.ptr(out),
                                     bh, nq, nk, d, _lib.stream_ptr()), "sb_attn_aggregate")
    return out


def aggregate(attn, fmap, to_v_weight, gamma, project_weight=None, heads: int = 1):
    """``Aggregate.forward`` (gma.py:102-115): ``fmap + gamma * project(attn @ v)``."""
    fm = _lib.dev_f32(fmap, "fmap")
    b, c, h, w = fm.shape
    n = h * w
    v = F.conv2d(fm, to_v_weight)
    inner = v.shape[1]
    d = inner // heads
    a = attn.reshape(b * heads, n, n)
    vv = v.reshape(b * heads, d, n)
    if project_weight is None and inner == c:
        out = attn_matmul_v(a, vv, residual=fm.reshape(b * heads, d, n), gamma=gamma)
        return out.view(b, c, h, w)
    out = attn_matmul_v(a, vv).view(b, inner, h, w)
    if project_weight is not None:
        out = F.conv2d(out, project_weight)
    return fm + gamma * out


def attention_forward(self, fmap):
    """Drop-in body for the reference's ``Attention.forward(self, fmap)`` (uses its ``to_qk``, ``heads``, ``scale``)."""
    return attention(fmap, self.to_qk.weight, self.heads, self.scale)


def aggregate_forward(self, attn, fmap):
    """Drop-in body for the reference's ``Aggregate.forward(self, attn, fmap)``."""
    return aggregate(attn, fmap, self.to_v.weight, self.gamma,
                     None if self.project is None else self.project.weight, self.heads)


class RelPosEmb(nn.Module):
    """State of the reference's ``RelPosEmb`` (gma.py:6-18): two embeddings and an index buffer.  The
    reference's ``Attention.forward`` never calls it (gma.py:54-76), so it has no forward here either; it
    exists so that ``Attention.state_dict()`` has the reference's keys (``pos_emb.rel_height.weight``,
    ``pos_emb.rel_width.weight``, ``pos_emb.rel_ind``) and its checkpoints load with ``strict=True``
    (evaluate.py:123, out.py:75,85)."""

    def __init__(self, max_pos_size, dim_head):
        super().__init__()
        self.rel_height = nn.Embedding(2 * max_pos_size - 1, dim_head)
        self.rel_width = nn.Embedding(2 * max_pos_size - 1, dim_head)
        deltas = torch.arange(max_pos_size).view(1, -1) - torch.arange(max_pos_size).view(-1, 1)
        self.register_buffer("rel_ind", deltas + max_pos_size - 1)


class Attention(nn.Module):
    """Same constructor, parameters and state_dict keys as the reference's ``Attention`` (gma.py:35-52)."""

    def __init__(self, *, args=None, dim, max_pos_size=100, heads=4, dim_head=128, attn_dtype=torch.float32):
        super().__init__()
        self.args, self.heads, self.scale = args, heads, dim_head ** -0.5
        self.to_qk = nn.Conv2d(dim, heads * dim_head * 2, 1, bias=False)
        self.pos_emb = RelPosEmb(max_pos_size, dim_head)
        self.attn_dtype = attn_dtype

    def forward(self, fmap):
        return attention(fmap, self.to_qk.weight, self.heads, self.scale, dtype=self.attn_dtype)


class Aggregate(nn.Module):
    """Same constructor / parameters as the reference's ``Aggregate`` (gma.py:79-100)."""

    def __init__(self, args=None, dim=128, heads=4, dim_head=128):
        super().__init__()
        self.args, self.heads, self.scale = args, heads, dim_head ** -0.5
        inner = heads * dim_head
        self.to_v = nn.Conv2d(dim, inner, 1, bias=False)
        self.gamma = nn.Parameter(torch.zeros(1))
        self.project = nn.Conv2d(inner, dim, 1, bias=False) if dim != inner else None

    def forward(self, attn, fmap):
        return aggregate(attn, fmap, self.to_v.weight, self.gamma,
                         None if self.project is None else self.project.weight, self.heads)
