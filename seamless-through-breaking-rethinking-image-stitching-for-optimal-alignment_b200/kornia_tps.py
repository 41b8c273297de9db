"""W3k — kornia-style thin-plate-spline image warp.

Mirrors the reference's ``core/inference/tps_methods/kornia_tps.py`` (the ``tps_method="kornia"``
branch of ``core/inference/tps_pipline.py:364-381``): ``get_tps_transform`` (vendored there as
``custom_get_tps_transform``, :47-103), ``warp_points_tps`` / ``create_meshgrid`` (imported there
from kornia; restated here from kornia's published source — kornia is absent from this image and
unpinned by the reference) and ``warp_image_tps`` (:105-176).  The small (K+3)^2 solve stays in
torch; the dense per-pixel evaluation + ``grid_sample`` is one fused CUDA kernel.
"""
from __future__ import annotations

import torch

from . import _lib

__all__ = ["get_tps_transform", "warp_points_tps", "warp_image_tps", "create_meshgrid", "grid_sample"]


def _pair_square_euclidean(t1, t2):
    """kornia_tps.py:26-36."""
    t1_sq = t1.mul(t1).sum(dim=-1, keepdim=True)
    t2_sq = t2.mul(t2).sum(dim=-1, keepdim=True).transpose(1, 2)
    t1_t2 = t1.matmul(t2.transpose(1, 2))
    return (-2 * t1_t2 + t1_sq + t2_sq).clamp(min=0)


def _kernel_distance(sd, eps=1e-8):
    """kornia_tps.py:38-45: 0.5 * r^2 * log(r^2 + eps)."""
    return 0.5 * sd * sd.add(eps).log()


def get_tps_transform(points_src, points_dst):
    """``(kernel_weights [B,N,2], affine_weights [B,3,2])`` warping src -> dst
    (custom_get_tps_transform, kornia_tps.py:47-103; pseudo-inverse solve :23)."""
    if points_src.dim() != 3 or points_dst.dim() != 3:
        raise ValueError(f"get_tps_transform: expected BxNx2 points, got {tuple(points_src.shape)} {tuple(points_dst.shape)}")
    device, dtype = points_src.device, points_src.dtype
    b, n = points_src.shape[:2]
    k = _kernel_distance(_pair_square_euclidean(points_src, points_dst))
    zero = torch.zeros(b, 3, 3, device=device, dtype=dtype)
    one = torch.ones(b, n, 1, device=device, dtype=dtype)
    dest = torch.cat((points_dst, zero[:, :, :2]), 1)
    p = torch.cat((one, points_src), -1)
    p_t = torch.cat((p, zero), 1).transpose(1, 2)
    l = torch.cat((torch.cat((k, p), -1), p_t), 1)
    sdt = dtype if dtype in (torch.float32, torch.float64) else torch.float32
    w = torch.pinverse(l.to(sdt)).matmul(dest.to(sdt)).to(dtype)
    return w[:, :-3], w[:, -3:]


def warp_points_tps(points_src, kernel_centers, kernel_weights, affine_weights):
    """kornia.geometry.transform.warp_points_tps (small point sets; plain torch)."""
    k = _kernel_distance(_pair_square_euclidean(points_src, kernel_centers))
    return (k[..., None].mul(kernel_weights[:, None]).sum(-2)
            + points_src[..., None].mul(affine_weights[:, None, 1:]).sum(-2)
            + affine_weights[:, None, 0])


def _axis_table(n, device):
    t = torch.linspace(0, n - 1, n, device=device, dtype=torch.float32)
    return ((t / (n - 1) - 0.5) * 2).contiguous()


def create_meshgrid(height, width, normalized_coordinates=True, device=None, dtype=torch.float32):
    """kornia.utils.create_meshgrid: ``[1, H, W, 2]`` (x, y)."""
    xs = torch.linspace(0, width - 1, width, device=device, dtype=dtype)
    ys = torch.linspace(0, height - 1, height, device=device, dtype=dtype)
    if normalized_coordinates:
        xs = (xs / (width - 1) - 0.5) * 2
        ys = (ys / (height - 1) - 0.5) * 2
    return torch.stack(torch.meshgrid([xs, ys], indexing="ij"), dim=-1).permute(1, 0, 2).unsqueeze(0)


def grid_sample(image, grid, align_corners=False):
    """``F.grid_sample(image, grid, mode='bilinear', padding_mode='zeros', align_corners)``."""
    lib = _lib.load()
    im = _lib.dev_f32(image, "image")
    gr = _lib.dev_f32(grid, "grid")
    n, c, h, w = im.shape
    if gr.dim() != 4 or gr.shape[0] != n or gr.shape[-1] != 2:
        raise ValueError(f"grid_sample: grid {tuple(gr.shape)} does not match image {tuple(im.shape)}")
    ho, wo = gr.shape[1], gr.shape[2]
    out = torch.empty((n, c, ho, wo), dtype=torch.float32, device=im.device)
    _lib.check(lib.sb_grid_sample(_lib.ptr(im), _lib.ptr(gr), _lib.ptr(out), n, c, h, w, ho, wo,
                                  1 if align_corners else 0, _lib.stream_ptr()), "sb_grid_sample")
    return out


def warp_image_tps(image, kernel_centers, kernel_weights, affine_weights, align_corners=False, return_grid=False):
    """image ``[B,C,H,W]``, kernel_centers / kernel_weights ``[B,K,2]``, affine_weights ``[B,3,2]``
    -> warped image ``[B,C,H,W]`` (kornia_tps.py:105-176)."""
    for name, t, nd in (("image", image, 4), ("kernel_centers", kernel_centers, 3),
                        ("kernel_weights", kernel_weights, 3), ("affine_weights", affine_weights, 3)):
        if not isinstance(t, torch.Tensor):
            raise TypeError(f"Input {name} is not torch.Tensor. Got {type(t)}")
        if t.dim() != nd:
            raise ValueError(f"Invalid shape for {name}. Got {tuple(t.shape)}")
    lib = _lib.load()
    im = _lib.dev_f32(image, "image")
    kc = _lib.dev_f32(kernel_centers, "kernel_centers")
    kw = _lib.dev_f32(kernel_weights, "kernel_weights")
    aw = _lib.dev_f32(affine_weights, "affine_weights")
    b, c, h, w = im.shape
    k = kc.shape[1]
    if kc.shape != (b, k, 2) or kw.shape != (b, k, 2) or aw.shape != (b, 3, 2):
        raise ValueError(f"warp_image_tps: inconsistent shapes {tuple(kc.shape)} {tuple(kw.shape)} {tuple(aw.shape)} for B={b}")
    xs, ys = _axis_table(w, im.device), _axis_table(h, im.device)
    out = torch.empty_like(im)
    grid = torch.empty((b, h, w, 2), dtype=torch.float32, device=im.device) if return_grid else None
    _lib.check(lib.sb_tps_kornia_warp(_lib.ptr(im), _lib.ptr(kc), _lib.ptr(kw), _lib.ptr(aw), _lib.ptr(xs),
                                      _lib.ptr(ys), _lib.ptr(out), _lib.ptr(grid), b, c, h, w, k,
                                      1 if align_corners else 0, _lib.stream_ptr()), "sb_tps_kornia_warp")
    return (out, grid) if return_grid else out
