"""W1 / W4 / G1 — flow warp, range-map occlusion and the small geometry helpers.

Mirrors the reference's ``core/warp_utils.py`` (same function names, argument
meaning and return shapes). ``warp`` and ``compute_range_map`` /
``compute_occlusion`` run as sm_100a kernels; the sub-millisecond geometry
helpers (``get_rigid_mesh``, ``H2Mesh``, ``resize_flow``) stay
plain torch, on whatever device their inputs live.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import _lib

__all__ = ["get_rigid_mesh", "H2Mesh", "resize_flow", "coords_grid", "flow_to_warp", "warp",
           "mask_invalid", "compute_range_map", "compute_fb_consistency", "compute_occlusion"]


# ------------------------------------------------------------------ G1 (torch)
def get_rigid_mesh(batch_size, height, width, grid_h=511, grid_w=511, device=None):
    """Regular (grid_h+1) x (grid_w+1) mesh of (x, y) points (warp_utils.py:10-18).
    The linspace tables are built on the CPU like the reference's, then moved."""
    xs = torch.linspace(0.0, float(width), grid_w + 1)
    ys = torch.linspace(0.0, float(height), grid_h + 1)
    if device is None:
        device = torch.device("cuda") if torch.cuda.is_available() else torch.device("cpu")
    xs, ys = xs.to(device), ys.to(device)
    ww = xs[None, :].expand(grid_h + 1, grid_w + 1)
    hh = ys[:, None].expand(grid_h + 1, grid_w + 1)
    mesh = torch.stack((ww, hh), dim=2)
    return mesh.unsqueeze(0).expand(batch_size, -1, -1, -1)


def H2Mesh(H, rigid_mesh, grid_h=511, grid_w=511):
    """Push the rigid mesh through H^-1 (warp_utils.py:20-34)."""
    h_inv = torch.inverse(H)
    b = rigid_mesh.shape[0]
    pts = rigid_mesh.reshape(b, -1, 2).to(H.device)
    ones = torch.ones(b, pts.shape[1], 1, device=H.device, dtype=pts.dtype)
    hom = torch.cat((pts, ones), dim=2)
    tar = torch.matmul(h_inv, hom.permute(0, 2, 1))
    mx = tar[:, 0, :] / tar[:, 2, :]
    my = tar[:, 1, :] / tar[:, 2, :]
    return torch.stack((mx, my), dim=2).reshape(b, grid_h + 1, grid_w + 1, 2)


def resize_flow(flow, new_shape):
    """Bilinear (align_corners=True) resize with per-axis rescale (warp_utils.py:38-46)."""
    _, _, h, w = flow.shape
    new_h, new_w = new_shape
    flow = F.interpolate(flow, (new_h, new_w), mode="bilinear", align_corners=True)
    scale_h, scale_w = h / float(new_h), w / float(new_w)
    flow[:, 0] /= scale_w
    flow[:, 1] /= scale_h
    return flow


def coords_grid(batch, ht, wd, device=None):
    ys, xs = torch.meshgrid(torch.arange(ht, device=device), torch.arange(wd, device=device), indexing="ij")
    return torch.stack([xs, ys], dim=0).float()[None].repeat(batch, 1, 1, 1)


def flow_to_warp(flow):
    """Flow end points ``[B,H,W,2]`` (warp_utils.py:54-69)."""
    b, _, h, w = flow.shape
    grid = coords_grid(b, h, w, device=flow.device).permute(0, 2, 3, 1)
    return grid + flow.permute(0, 2, 3, 1)


def mask_invalid(coords, pad_h=0, pad_w=0):
    """1 where coords ``[B,H,W,2]`` lie inside the image (warp_utils.py:83-111)."""
    if coords.dim() != 4:
        raise NotImplementedError()
    max_h = float(coords.shape[-3] - 1)
    max_w = float(coords.shape[-2] - 1)
    m = (coords[..., 0] >= float(pad_w)) & (coords[..., 0] <= max_w) & \
        (coords[..., 1] >= float(pad_h)) & (coords[..., 1] <= max_h)
    return m.float()[:, None]


# ------------------------------------------------------------------ W1 (kernel)
def warp(x, flo, mode="bilinear", mul_mask=None, return_overlap=False):
    """Backward warp of ``x [B,C,H,W]`` by ``flo [B,2,H,W]`` (warp_utils.py:71-80).

    ``mode='nearest'`` follows the reference to the letter: its non-bilinear branch calls ``F.grid_sample`` without
    ``align_corners`` (:79), so the nearest mode samples with ``align_corners=False``.

    ``mul_mask`` ([B,1,H,W], optional, not in the reference signature) fuses the
    caller's ``final_warp_output * mask`` (flowHomoAdpater.py:182,317) into the
    same pass; ``return_overlap=True`` (C == 6) also returns
    ``where(mean(out[:,3:6]) < 0.9, 1, 0)`` of the unmasked warp (:171-174)."""
    if mode not in ("bilinear", "nearest"):
        raise ValueError(f"warp: mode must be 'bilinear' or 'nearest' (F.grid_sample's modes the reference passes on), got {mode!r}")
    lib = _lib.load()
    if mode == "nearest":
        if mul_mask is not None or return_overlap:
            raise ValueError("warp: mul_mask / return_overlap are extensions of the bilinear mode")
        xs = _lib.dev_f32(x, "x")
        fl = _lib.dev_f32(flo, "flo")
        if xs.dim() != 4 or fl.dim() != 4 or fl.shape[1] != 2 or xs.shape[0] != fl.shape[0] or xs.shape[2:] != fl.shape[2:]:
            raise ValueError(f"warp: x {tuple(xs.shape)} and flo {tuple(fl.shape)} do not match")
        b, c, h, w = xs.shape
        out = torch.empty_like(xs)
        _lib.check(lib.sb_flow_warp_nearest(_lib.ptr(xs), _lib.ptr(fl), _lib.ptr(out), b, c, h, w, _lib.stream_ptr()),
                   "sb_flow_warp_nearest")
        return out
    xs = _lib.dev_f32(x, "x")
    fl = _lib.dev_f32(flo, "flo")
    if xs.dim() != 4 or fl.dim() != 4 or fl.shape[1] != 2 or xs.shape[0] != fl.shape[0] or xs.shape[2:] != fl.shape[2:]:
        raise ValueError(f"warp: x {tuple(xs.shape)} and flo {tuple(fl.shape)} do not match")
    b, c, h, w = xs.shape
    mm = None
    if mul_mask is not None:
        mm = _lib.dev_f32(mul_mask, "mul_mask")
        if mm.numel() != b * h * w:
            raise ValueError("warp: mul_mask must be [B,1,H,W]")
    out = torch.empty_like(xs)
    ov = torch.empty((b, h, w), dtype=torch.float32, device=xs.device) if return_overlap else None
    _lib.check(lib.sb_flow_warp(_lib.ptr(xs), _lib.ptr(fl), _lib.ptr(mm), _lib.ptr(out), _lib.ptr(ov),
                                b, c, h, w, _lib.stream_ptr()), "sb_flow_warp")
    return (out, ov) if return_overlap else out


# ------------------------------------------------------------------ W4 (kernel)
def _range_map(flow, mode):
    lib = _lib.load()
    fl = _lib.dev_f32(flow, "flow")
    if fl.dim() != 4 or fl.shape[1] != 2:
        raise ValueError(f"compute_range_map: flow must be [B,2,H,W], got {tuple(fl.shape)}")
    b, _, h, w = fl.shape
    out = torch.empty((b, 1, h, w), dtype=torch.float32, device=fl.device)
    accum = torch.empty((b, h, w), dtype=torch.int64, device=fl.device)
    _lib.check(lib.sb_range_map(_lib.ptr(fl), _lib.ptr(accum), _lib.ptr(out), b, h, w, mode,
                                _lib.stream_ptr()), "sb_range_map")
    return out


def compute_range_map(flow):
    """Forward-splat count of the backward flow (warp_utils.py:114-175).
    Deterministic (fixed-point integer atomics) where the reference's
    ``scatter_add_`` is order-dependent on a GPU."""
    return _range_map(flow, 0)


def compute_fb_consistency(flow_ij, flow_ji):
    """warp_utils.py:177-183."""
    flow_ji_in_i = warp(flow_ji, flow_ij)
    fb_sq_diff = torch.sum((flow_ij + flow_ji_in_i) ** 2, dim=1, keepdim=True)
    fb_sum_sq = torch.sum((flow_ij ** 2 + flow_ji_in_i ** 2), dim=1, keepdim=True)
    return fb_sq_diff, fb_sum_sq


def compute_occlusion(flow_ij, flow_ji, occlusion_estimation, occlusion_are_zeros=False,
                      boundaries_occluded=True, threshold=False):
    """Occlusion mask ``[B,1,H,W]`` (warp_utils.py:185-221).

    The 'wang' estimator (the one every call site uses) is a single fused
    splat + finalise; the forward/backward warp the reference computes and then
    discards for 'wang' is skipped. ``threshold=True`` (extension) also fuses the
    caller's ``>= 0.5`` binarisation (flowHomoAdpater.py:181)."""
    if occlusion_estimation == "wang":
        if occlusion_are_zeros and boundaries_occluded:
            return _range_map(flow_ji, 3 if threshold else 1)
        occlusion_mask = _range_map(flow_ji, 2)
    elif occlusion_estimation == "none":
        occlusion_mask = torch.zeros_like(flow_ij[:, :1])
    elif occlusion_estimation in ("brox", "fb_abs"):
        fb_sq_diff, fb_sum_sq = compute_fb_consistency(flow_ij, flow_ji)
        if occlusion_estimation == "brox":
            occlusion_mask = (fb_sq_diff > 0.01 * fb_sum_sq + 0.5).float()
        else:
            occlusion_mask = (fb_sq_diff ** 0.5 > 1.5).float()
    else:
        occlusion_mask = torch.zeros_like(flow_ij[:, :1])
    if not boundaries_occluded:
        occlusion_mask = torch.min(occlusion_mask, mask_invalid(flow_to_warp(flow_ij)))
    if occlusion_are_zeros:
        occlusion_mask = 1 - occlusion_mask
    if threshold:
        occlusion_mask = (occlusion_mask >= 0.5).float()
    return occlusion_mask
