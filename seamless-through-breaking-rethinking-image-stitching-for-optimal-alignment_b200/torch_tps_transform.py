"""W3 — thin-plate-spline backward warp with the UDIS sampler.

Mirrors ``transformer(U, source, target, out_size)`` of the reference's
``core/udis_utils/torch_tps_transform.py:7-191`` (``torch_tps_transform2.py`` is
the same maths). The (pn+3)^2 fp64 solve stays in torch; the dense basis
evaluation + sampling is one fused kernel (the reference materialises a
``[B, pn+3, H*W]`` tensor)."""
from __future__ import annotations

import torch

from . import _lib
from .torch_homo_transform import _as_int, linspace_table

__all__ = ["transformer", "solve_system"]


def solve_system(source, target):
    """TPS coefficients ``T [B,2,pn+3]`` (torch_tps_transform.py:149-185):
    W = [[P, K], [0, P^T]], K = d2 * log(d2 + 1e-6), fp64 inverse."""
    b, pn = source.shape[0], source.shape[1]
    dev = source.device
    source = source.float()
    ones = torch.ones(b, pn, 1, device=dev)
    p = torch.cat([ones, source], 2)
    d2 = torch.sum(torch.square(p.reshape(b, -1, 1, 3) - p.reshape(b, 1, -1, 3)), 3)
    r = d2 * torch.log(d2 + 1e-6)
    w0 = torch.cat((p, r), 2)
    w1 = torch.cat((torch.zeros(b, 3, 3, device=dev), p.permute(0, 2, 1)), 2)
    w = torch.cat((w0, w1), 1)
    w_inv = torch.inverse(w.double())
    tp = torch.cat((target.float(), torch.zeros(b, 3, 2, device=dev)), 1)
    t = torch.matmul(w_inv, tp.double())
    return t.permute(0, 2, 1).float().contiguous()


def transformer(U, source, target, out_size, return_indices=False):
    """U ``[B,C,H,W]``; source, target ``[B,pn,2]`` control points in [-1,1];
    warps U from ``target`` to ``source`` -> ``[B,C,Hout,Wout]``."""
    lib = _lib.load()
    u = _lib.dev_f32(U, "U")
    src = _lib.dev_f32(source, "source")
    tgt = _lib.dev_f32(target, "target")
    b, c, h, w = u.shape
    if src.shape != tgt.shape or src.dim() != 3 or src.shape[0] != b or src.shape[2] != 2:
        raise ValueError(f"transformer: control points {tuple(src.shape)} / {tuple(tgt.shape)} do not match B={b}")
    pn = src.shape[1]
    T = solve_system(src, tgt)
    hout, wout = _as_int(out_size[0]), _as_int(out_size[1])
    xs, ys = linspace_table(wout, u.device), linspace_table(hout, u.device)
    out = torch.empty((b, c, hout, wout), dtype=torch.float32, device=u.device)
    idx = torch.empty((b, 4, hout, wout), dtype=torch.int32, device=u.device) if return_indices else None
    _lib.check(lib.sb_tps_warp(_lib.ptr(u), _lib.ptr(T), _lib.ptr(src), _lib.ptr(xs), _lib.ptr(ys),
                               _lib.ptr(out), _lib.ptr(idx), b, c, h, w, hout, wout, pn, _lib.stream_ptr()),
               "sb_tps_warp")
    return (out, idx) if return_indices else out
