"""G1 — 4-point DLT (reference ``core/udis_utils/torch_DLT.py:17-45``).

Eight unknowns of H from four correspondences; a batched 8x8 solve, tiny, kept
in torch (library LU) on the inputs' device."""
from __future__ import annotations

import torch

__all__ = ["tensor_DLT"]


def tensor_DLT(src_p, dst_p):
    bs = src_p.shape[0]
    dev, dt = src_p.device, src_p.dtype
    ones = torch.ones(bs, 4, 1, device=dev, dtype=dt)
    xy1 = torch.cat((src_p, ones), 2)
    zeros = torch.zeros_like(xy1)
    # rows alternate (x y 1 0 0 0) / (0 0 0 x y 1)
    m1 = torch.cat((torch.cat((xy1, zeros), 2), torch.cat((zeros, xy1), 2)), 2).reshape(bs, -1, 6)
    m2 = torch.matmul(dst_p.reshape(-1, 2, 1), src_p.reshape(-1, 1, 2)).reshape(bs, -1, 2)
    a = torch.cat((m1, -m2), 2)
    rhs = dst_p.reshape(bs, -1, 1)
    h8 = torch.matmul(torch.inverse(a), rhs).reshape(bs, 8)
    return torch.cat((h8, ones[:, 0, :]), 1).reshape(bs, 3, 3)
