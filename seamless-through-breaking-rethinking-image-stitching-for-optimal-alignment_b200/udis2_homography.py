"""N3 — contextual correlation layer of the UDIS2 homography network ("next" row 3 of SURVEY §8f).

Mirrors ``UDIS2Network.CCL(self, feature_1, feature_2)`` of the reference
(``core/UDIS2/Homography/network.py:147-199``, called at ``:130`` on the 1/16-resolution ResNet
features ``[B, 1024, 32, 32]``).  See ``csrc/ccl.cu`` for the restructuring (plain all-pairs
correlation on the TF32 tensor cores + nine shifted diagonals + online softmax expectation).
"""
from __future__ import annotations

import torch

from . import _lib

__all__ = ["CCL", "udis2_network_ccl", "gemm_nt_tf32"]

_ws = {}


def _workspace(nbytes, device):
    key = (device.type, device.index)
    w = _ws.get(key)
    if w is None or w.numel() < nbytes:
        w = torch.empty(nbytes + 256, dtype=torch.uint8, device=device)
        _ws[key] = w
    off = (-w.data_ptr()) % 256
    return w[off:off + nbytes]


def CCL(feature_1, feature_2, softmax_scale: float = 10.0):
    """feature_1, feature_2 ``[B,C,H,W]`` -> feature flow ``[B,2,H,W]`` (channel 0 = w, 1 = h)."""
    lib = _lib.load()
    f1 = _lib.dev_f32(feature_1, "feature_1")
    f2 = _lib.dev_f32(feature_2, "feature_2")
    if f1.dim() != 4 or f1.shape != f2.shape:
        raise ValueError(f"CCL: expected equal [B,C,H,W] features, got {tuple(f1.shape)} {tuple(f2.shape)}")
    b, c, h, w = f1.shape
    out = torch.empty((b, 2, h, w), dtype=torch.float32, device=f1.device)
    if out.numel() == 0:
        return out
    need = lib.sb_ccl_workspace_bytes(b, c, h, w)
    ws = _workspace(need, f1.device)
    _lib.check(lib.sb_ccl(_lib.ptr(f1), _lib.ptr(f2), _lib.ptr(out), _lib.ptr(ws), ws.numel(), b, c, h, w,
                          float(softmax_scale), _lib.stream_ptr()), "sb_ccl")
    return out


def udis2_network_ccl(self, feature_1, feature_2):
    """Drop-in body for ``UDIS2Network.CCL(self, feature_1, feature_2)``."""
    return CCL(feature_1, feature_2)


def gemm_nt_tf32(a, b):
    """``a [BH,M,K] @ b [BH,N,K]^T -> [BH,M,N]`` on the TF32 tensor cores (the kernel CCL and GMA share)."""
    lib = _lib.load()
    a = _lib.dev_f32(a, "a")
    b = _lib.dev_f32(b, "b")
    bh, m, k = a.shape
    n = b.shape[1]
    if b.shape != (bh, n, k):
        raise ValueError(f"gemm_nt_tf32: shapes {tuple(a.shape)} {tuple(b.shape)}")
    d = torch.empty((bh, m, n), dtype=torch.float32, device=a.device)
    _lib.check(lib.sb_gemm_nt_tf32(_lib.ptr(a), _lib.ptr(b), _lib.ptr(d), bh, m, n, k, _lib.stream_ptr()), "sb_gemm_nt_tf32")
    return d
