"""G1 — 4-point DLT (reference ``core/udis_utils/torch_DLT.py:17-45``) and the
normalised-coordinate homographies the adapter derives from it
(``core/flowHomoAdpater.py:96-113``).

On a GPU both run as ONE fused launch (``sb_dlt_theta``): the reference's path is
~40 tiny ATen launches plus ``torch.inverse``, whose singularity check
synchronises the host on every call."""
from __future__ import annotations

import ctypes

import torch

from . import _lib

__all__ = ["tensor_DLT", "dlt_thetas", "norm_matrix", "corner_points"]

_IDENT = (ctypes.c_float * 9)(1, 0, 0, 0, 1, 0, 0, 0, 1)


def norm_matrix(w, h):
    """M maps normalised [-1,1] coordinates to pixels of a w x h image (row-major list)."""
    return [w / 2.0, 0.0, w / 2.0, 0.0, h / 2.0, h / 2.0, 0.0, 0.0, 1.0]


def _inv3(m):
    import numpy as np
    return np.linalg.inv(np.asarray(m, dtype=np.float64).reshape(3, 3)).reshape(-1).tolist()


def dlt_thetas(src_p, dst_p, left=None, right=None, want_inverse=True):
    """H = DLT(src_p, dst_p); theta = left @ H @ right; theta_inv = left @ H^-1 @ right.
    ``left`` / ``right``: row-major 3x3 as 9 python floats (host constants; default identity).
    Returns (H, theta, theta_inv) as ``[B,3,3]`` fp32 CUDA tensors."""
    lib = _lib.load()
    s = _lib.dev_f32(src_p, "src_p")
    d = _lib.dev_f32(dst_p, "dst_p")
    if s.shape != d.shape or s.dim() != 3 or s.shape[1:] != (4, 2):
        raise ValueError(f"dlt_thetas: expected [B,4,2] point sets, got {tuple(s.shape)} {tuple(d.shape)}")
    b = s.shape[0]
    L = (ctypes.c_float * 9)(*left) if left is not None else _IDENT
    R = (ctypes.c_float * 9)(*right) if right is not None else _IDENT
    H = torch.empty((b, 3, 3), dtype=torch.float32, device=s.device)
    theta = torch.empty_like(H)
    theta_inv = torch.empty_like(H) if want_inverse else None
    _lib.check(lib.sb_dlt_theta(_lib.ptr(s), _lib.ptr(d), ctypes.cast(L, ctypes.c_void_p),
                                ctypes.cast(R, ctypes.c_void_p), _lib.ptr(H), _lib.ptr(theta),
                                _lib.ptr(theta_inv), b, _lib.stream_ptr()), "sb_dlt_theta")
    return H, theta, theta_inv


def tensor_DLT(src_p, dst_p):
    """Eight unknowns of H from four correspondences -> ``[B,3,3]`` with H[2,2] = 1."""
    return dlt_thetas(src_p, dst_p, want_inverse=False)[0]


_corner_cache: dict = {}


def corner_points(w, h, batch, device):
    """The four source corners [[0,0],[w,0],[0,h],[w,h]] as a cached ``[B,4,2]`` device
    constant (building it per call would cost a synchronous host->device copy)."""
    key = (float(w), float(h), int(batch), str(device))
    t = _corner_cache.get(key)
    if t is None:
        t = torch.tensor([[0.0, 0.0], [w, 0.0], [0.0, h], [w, h]], dtype=torch.float32)
        t = t.unsqueeze(0).repeat(batch, 1, 1).to(device)
        _corner_cache[key] = t
    return t
