"""N4 — the cost-map PatchEmbed projection ("next" row 4 of SURVEY §8f).

Mirrors the convolution stack of the reference's ``PatchEmbed``
(``core/FlowFormer/PerCostFormer3/encoder.py:20-92``: built at ``:36-43`` for ``patch_embed = "single"``,
``patch_size = 8``; run at ``:68-73``; applied to every cost map at ``:263``):

    Conv2d(1, 16, 6, stride 2, pad 2) -> ReLU -> Conv2d(16, 32, 6, 2, 2) -> ReLU -> Conv2d(32, 64, 6, 2, 2)

over ``cost_maps [B*H1*W1, 1, 64, 64]`` — 20 MFLOP per map, 1.3 TFLOP per direction at batch 16, and the
only other consumer that streams the whole fp32 cost volume.  One tcgen05 kernel (CTA pairs,
``csrc/patch_embed.cu``) keeps all three layers' activations in shared memory.  bf16 operands with fp32
accumulation; the position encoding / 1x1 FFN / LayerNorm that follow (``:75-92``) are network code and stay
what the reference runs.
"""
from __future__ import annotations

import sys
import weakref

import torch
import torch.nn.functional as F

from . import _lib

__all__ = ["patch_embed_proj", "pack_patch_embed_weights", "patch_embed_forward", "PatchEmbedProj"]


def pack_patch_embed_weights(w1, w2, w3):
    """fp32 conv weights ``[16,1,6,6]``, ``[32,16,6,6]``, ``[64,32,6,6]`` -> the kernel's packed bf16 images
    (a ``uint8`` CUDA tensor; reuse it for every call with the same weights)."""
    lib = _lib.load()
    ws = [_lib.dev_f32(w, f"w{i + 1}") for i, w in enumerate((w1, w2, w3))]
    if tuple(ws[0].shape) != (16, 1, 6, 6) or tuple(ws[1].shape) != (32, 16, 6, 6) or tuple(ws[2].shape) != (64, 32, 6, 6):
        raise ValueError("pack_patch_embed_weights: expected conv weights [16,1,6,6], [32,16,6,6], [64,32,6,6] "
                         f"(embed_dim = 64, patch_size = 8), got {[tuple(w.shape) for w in ws]}")
    pack = torch.empty(lib.sb_patch_embed_pack_bytes(), dtype=torch.uint8, device=ws[0].device)
    _lib.check(lib.sb_patch_embed_pack(_lib.ptr(ws[0]), _lib.ptr(ws[1]), _lib.ptr(ws[2]), _lib.ptr(pack),
                                       _lib.stream_ptr()), "sb_patch_embed_pack")
    return pack


def patch_embed_proj(x, w1, b1, w2, b2, w3, b3, pack=None):
    """x ``[N,1,64,64]`` -> ``[N,64,8,8]``: the three strided convolutions with the two ReLUs between them.
    ``pack``: result of :func:`pack_patch_embed_weights` for these weights (made on the fly when None)."""
    lib = _lib.load()
    xs = _lib.dev_f32(x, "x")
    if xs.dim() != 4 or xs.shape[1] != 1:
        raise ValueError(f"patch_embed_proj: x must be [N,1,H,W], got {tuple(xs.shape)}")
    n, _, h, w = xs.shape
    if (h, w) != (64, 64):
        raise NotImplementedError(f"patch_embed_proj: {h}x{w} cost maps; the kernel is specialised for 64x64 "
                                  "(512x512 images, the shipped configuration)")
    if pack is None:
        pack = pack_patch_embed_weights(w1, w2, w3)
    bias = torch.cat([_lib.dev_f32(b, "bias").reshape(-1) for b in (b1, b2, b3)])
    if bias.numel() != 112:
        raise ValueError("patch_embed_proj: biases must have 16, 32 and 64 elements")
    out = torch.empty((n, 64, 8, 8), dtype=torch.float32, device=xs.device)
    _lib.check(lib.sb_patch_embed_proj(_lib.ptr(xs), _lib.ptr(pack), _lib.ptr(bias), _lib.ptr(out), n, h, w,
                                       _lib.stream_ptr()), "sb_patch_embed_proj")
    return out


_packs = weakref.WeakKeyDictionary()     # proj ModuleList -> (weight versions, pack)


def _module_pack(proj):
    convs = (proj[0], proj[2], proj[4])
    key = tuple((c.weight.data_ptr(), c.weight._version, str(c.weight.device)) for c in convs)
    hit = _packs.get(proj)
    if hit is None or hit[0] != key:
        hit = (key, pack_patch_embed_weights(*(c.weight.detach() for c in convs)))
        _packs[proj] = hit
    return hit[1]


def _fusable(self, x, masks):
    p = self.proj
    return (all(m is None for m in masks) and isinstance(p, torch.nn.ModuleList) and len(p) == 5
            and x.dim() == 4 and tuple(x.shape[1:]) == (1, 64, 64) and self.patch_size == 8
            and isinstance(p[0], torch.nn.Conv2d) and p[0].out_channels == 16 and p[2].out_channels == 32
            and p[4].out_channels == 64 and p[0].bias is not None)


def patch_embed_forward(self, x, mask_for_patch1=None, mask_for_patch2=None, mask_for_patch3=None):
    """Drop-in body for the reference's ``PatchEmbed.forward`` (encoder.py:60-92).  The convolution stack runs as
    the fused kernel whenever it is the shipped configuration on 64x64 maps without pre-training masks; anything
    else runs the module's own layers exactly as the reference does.  The tail (:75-92) is the reference's."""
    ref = sys.modules[type(self).__module__]          # coords_grid / position encodings of the reference module
    B, C, H, W = x.shape
    pad_r = (self.patch_size - W % self.patch_size) % self.patch_size
    pad_b = (self.patch_size - H % self.patch_size) % self.patch_size
    x = F.pad(x, (0, pad_r, 0, pad_b))
    masks = [mask_for_patch1, mask_for_patch2, mask_for_patch3]
    if _fusable(self, x, masks):
        p = self.proj
        x = patch_embed_proj(x, p[0].weight, p[0].bias, p[2].weight, p[2].bias, p[4].weight, p[4].bias,
                             pack=_module_pack(p))
    else:
        for idx, layer in enumerate(self.proj):
            if idx % 2 == 0 and masks[idx // 2] is not None:
                x = x * (1 - masks[idx // 2])
            x = layer(x)
    out_size = x.shape[2:]
    patch_coord = ref.coords_grid(B, out_size[0], out_size[1]).to(x.device) * self.patch_size + self.patch_size / 2
    if self.cfg.use_rpe:
        center_coord = ref.coords_grid(1, H, W).to(x.device)
        center_coord = center_coord.permute(2, 3, 1, 0).reshape(H * W, 2, 1, 1).repeat(B // (H * W), 1, 1, 1)
        patch_coord = patch_coord - center_coord
    patch_coord = patch_coord.view(B, 2, -1).permute(0, 2, 1)
    if self.pe == "linear":
        patch_coord_enc = ref.LinearPositionEmbeddingSine(patch_coord, dim=64)
    elif self.pe == "exp":
        patch_coord_enc = ref.ExpPositionEmbeddingSine(patch_coord, dim=64)
    patch_coord_enc = patch_coord_enc.permute(0, 2, 1).view(B, -1, out_size[0], out_size[1])
    x_pe = torch.cat([x, patch_coord_enc], dim=1)
    x = self.ffn_with_coord(x_pe)
    x = self.norm(x.flatten(2).transpose(1, 2))
    return x, out_size


class PatchEmbedProj(torch.nn.Module):
    """The convolution stack alone, with the reference's parameter names (``proj.0/2/4.weight|bias``)."""

    def __init__(self, in_chans=1, embed_dim=64):
        super().__init__()
        if in_chans != 1 or embed_dim != 64:
            raise NotImplementedError("PatchEmbedProj: the kernel is specialised for in_chans = 1, embed_dim = 64")
        self.proj = torch.nn.ModuleList([
            torch.nn.Conv2d(in_chans, embed_dim // 4, kernel_size=6, stride=2, padding=2), torch.nn.ReLU(),
            torch.nn.Conv2d(embed_dim // 4, embed_dim // 2, kernel_size=6, stride=2, padding=2), torch.nn.ReLU(),
            torch.nn.Conv2d(embed_dim // 2, embed_dim, kernel_size=6, stride=2, padding=2)])

    def forward(self, x):
        p = self.proj
        return patch_embed_proj(x, p[0].weight, p[0].bias, p[2].weight, p[2].bias, p[4].weight, p[4].bias,
                                pack=_module_pack(p))
