"""W2 — homography backward warp with the UDIS sampler.

Mirrors ``transformer(U, theta, out_size, **kwargs)`` of the reference's
``core/udis_utils/torch_homo_transform.py:5-151``."""
from __future__ import annotations

import torch

from . import _lib

__all__ = ["transformer", "linspace_table"]

_tables: dict = {}


def linspace_table(n: int, device) -> torch.Tensor:
    """``torch.linspace(-1, 1, n)`` evaluated on the CPU exactly as the reference
    does (torch_homo_transform.py:96-99), cached per (n, device). It is passed to
    the kernel as a table because linspace is not reproducible arithmetically."""
    key = (int(n), str(device))
    t = _tables.get(key)
    if t is None:
        t = torch.linspace(-1.0, 1.0, int(n)).to(device)
        _tables[key] = t
    return t


def _as_int(v) -> int:
    return int(v.item()) if isinstance(v, torch.Tensor) else int(v)


def transformer(U, theta, out_size, return_indices=False, append_ones=0, **kwargs):
    """U ``[B,C,H,W]``, theta ``[B or 1,3,3]`` (normalised [-1,1] coordinates),
    ``out_size=(Hout,Wout)`` ints or 0-dim tensors -> ``[B,C,Hout,Wout]``.

    ``return_indices=True`` (extension for the parity tests) also returns the
    clamped integer grid indices ``[B,4,Hout,Wout]`` int32 = (x0,x1,y0,y1).
    ``append_ones=n`` (extension) behaves exactly like
    ``transformer(torch.cat((U, ones[:, :n]), 1), ...)`` — what every call site of the
    reference does — without materialising or reading the ones planes."""
    lib = _lib.load()
    u = _lib.dev_f32(U, "U")
    if u.dim() != 4:
        raise ValueError(f"transformer: U must be [B,C,H,W], got {tuple(u.shape)}")
    th = _lib.dev_f32(theta, "theta").reshape(-1, 3, 3)
    b, c, h, w = u.shape
    if th.shape[0] not in (1, b):
        raise ValueError(f"transformer: theta batch {th.shape[0]} incompatible with B={b}")
    hout, wout = _as_int(out_size[0]), _as_int(out_size[1])
    xs, ys = linspace_table(wout, u.device), linspace_table(hout, u.device)
    n1 = int(append_ones)
    out = torch.empty((b, c + n1, hout, wout), dtype=torch.float32, device=u.device)
    idx = torch.empty((b, 4, hout, wout), dtype=torch.int32, device=u.device) if return_indices else None
    _lib.check(lib.sb_homo_warp(_lib.ptr(u), _lib.ptr(th), _lib.ptr(xs), _lib.ptr(ys), _lib.ptr(out),
                                _lib.ptr(idx), b, c, n1, h, w, hout, wout, th.shape[0], _lib.stream_ptr()),
               "sb_homo_warp")
    return (out, idx) if return_indices else out
