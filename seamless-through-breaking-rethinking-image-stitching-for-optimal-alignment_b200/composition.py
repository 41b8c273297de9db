"""W5 / W6 / W7 / W8 — mask morphology and fused overlap-mask compositing.

Mirrors ``preprocess_occlusion_mask`` (reference ``core/flowHomoAdpater.py:18-35``),
the inline compositing of ``test_out_forward`` (``:317,337-360``),
``build_model`` (``core/UDIS2/Composition/network.py:8-20``) and the TPS-stage
mix / blend of ``core/inference/tps_pipline.py:139-170``."""
from __future__ import annotations

import torch

from . import _lib

__all__ = ["preprocess_occlusion_mask", "morph_open", "composite_test_out", "build_model",
           "tps_mix_blend", "overlap_mask"]


def morph_open(mask, kernel_size=(19, 19), border_is_zero=True):
    """Binarise (>= 0.5) then erode + dilate with a ``kh x kw`` box. ``mask`` is
    ``[B,1,H,W]`` (or any ``[..., H, W]``); exact {0,1} fp32 output."""
    lib = _lib.load()
    m = _lib.dev_f32(mask, "mask")
    if m.dim() < 2:
        raise ValueError("morph_open: mask needs at least [H,W]")
    h, w = m.shape[-2:]
    planes = m.numel() // max(h * w, 1) if h * w else 0
    out = torch.empty_like(m)
    kh, kw = int(kernel_size[0]), int(kernel_size[1])
    _lib.check(lib.sb_morph_open(_lib.ptr(m), _lib.ptr(out), planes, h, w, kh, kw, 1 if border_is_zero else 0,
                                 _lib.stream_ptr()), "sb_morph_open")
    return out


def preprocess_occlusion_mask(occlusion_mask, kernel_size=(19, 19)):
    """flowHomoAdpater.py:18-35. Like the reference this is only meaningful for
    C == 1 (its ``kernel.numel()`` threshold counts channels)."""
    if occlusion_mask.dim() != 4 or occlusion_mask.shape[1] != 1:
        raise ValueError("preprocess_occlusion_mask: expected [B,1,H,W] (the reference's threshold "
                         "kernel.numel() is only correct for one channel)")
    return morph_open(occlusion_mask, kernel_size, border_is_zero=True)


def composite_test_out(homo_output, homo_output2, final_warp_in, occlusion_mask=None):
    """Fused compositing of ``test_out_forward`` (flowHomoAdpater.py:337-360).

    homo_output ``[B,6,h,w]`` (img1 | mask on the canvas), homo_output2 ``[B,6,h,w]``
    (H-warped img2 | mask), final_warp_in ``[B,6,h,w]`` (flow warp of homo_output2,
    already multiplied by the flow mask, :317), occlusion_mask ``[B,1,h,w]`` or None.
    Returns the ``out_dict`` entries this arithmetic produces."""
    lib = _lib.load()
    h1 = _lib.dev_f32(homo_output, "homo_output")
    h2 = _lib.dev_f32(homo_output2, "homo_output2")
    fw = _lib.dev_f32(final_warp_in, "final_warp_in")
    if not (h1.shape == h2.shape == fw.shape) or h1.dim() != 4 or h1.shape[1] != 6:
        raise ValueError("composite_test_out: the three inputs must all be [B,6,h,w]")
    b, _, h, w = h1.shape
    occ = None
    if occlusion_mask is not None:
        occ = _lib.dev_f32(occlusion_mask, "occlusion_mask")
        if occ.numel() != b * h * w:
            raise ValueError("composite_test_out: occlusion_mask must be [B,1,h,w]")
    dev = h1.device
    final_warp = torch.empty_like(fw)
    output2 = torch.empty((b, 3, h, w), dtype=torch.float32, device=dev)
    mask1 = torch.empty_like(output2)
    mask2 = torch.empty_like(output2)
    blend = torch.empty((b, 3, h, w), dtype=torch.uint8, device=dev)
    _lib.check(lib.sb_composite_test_out(_lib.ptr(h1), _lib.ptr(h2), _lib.ptr(fw), _lib.ptr(occ),
                                         _lib.ptr(final_warp), _lib.ptr(output2), _lib.ptr(mask1),
                                         _lib.ptr(mask2), _lib.ptr(blend), b, h, w, _lib.stream_ptr()),
               "sb_composite_test_out")
    return dict(final_warp_output=final_warp, output1=h1[:, 0:3], output2=output2, mask1=mask1, mask2=mask2,
                blend_image=blend)


def build_model(net, warp1_tensor, warp2_tensor, mask1_tensor, mask2_tensor):
    """Composition/network.py:8-20 — ``net`` (the UNet, not ours) predicts the seam
    mask; the learned-mask / stitched-image arithmetic is one fused kernel."""
    lib = _lib.load()
    out = net(warp1_tensor, warp2_tensor, mask1_tensor, mask2_tensor)
    w1 = _lib.dev_f32(warp1_tensor, "warp1_tensor")
    w2 = _lib.dev_f32(warp2_tensor, "warp2_tensor")
    m1 = _lib.dev_f32(mask1_tensor, "mask1_tensor")
    m2 = _lib.dev_f32(mask2_tensor, "mask2_tensor")
    o = _lib.dev_f32(out, "net output")
    b, c, h, w = w1.shape
    if c != 3 or not (w2.shape == m1.shape == m2.shape == w1.shape) or o.numel() != b * h * w:
        raise ValueError("build_model: expected four [B,3,h,w] tensors and a [B,1,h,w] net output")
    lm1, lm2, st = torch.empty_like(w1), torch.empty_like(w1), torch.empty_like(w1)
    _lib.check(lib.sb_build_model(_lib.ptr(w1), _lib.ptr(w2), _lib.ptr(m1), _lib.ptr(m2), _lib.ptr(o),
                                  _lib.ptr(lm1), _lib.ptr(lm2), _lib.ptr(st), b, h, w, _lib.stream_ptr()),
               "sb_build_model")
    return dict(learned_mask1=lm1, learned_mask2=lm2, stitched_image=st)


def tps_mix_blend(final_warp, tps_H_warp, tps_H_warp_mask, output1, mask1):
    """TPS-stage mix + average blend (tps_pipline.py:150-170).

    final_warp, output1, mask1 ``[B,3,h,w]``; tps_H_warp ``[B,3,h,w]`` already
    multiplied by tps_H_warp_mask ``[B,1,h,w]`` (the opened mask of :139-148, see
    :func:`tps_warp_mask`). Returns (output2, mask2 ``[B,1,h,w]``, blend uint8)."""
    lib = _lib.load()
    fw = _lib.dev_f32(final_warp, "final_warp")
    tw = _lib.dev_f32(tps_H_warp, "tps_H_warp")
    tm = _lib.dev_f32(tps_H_warp_mask, "tps_H_warp_mask")
    o1 = _lib.dev_f32(output1, "output1")
    m1 = _lib.dev_f32(mask1, "mask1")
    b, c, h, w = fw.shape
    if c != 3 or not (tw.shape == o1.shape == m1.shape == fw.shape) or tm.numel() != b * h * w:
        raise ValueError("tps_mix_blend: shape mismatch")
    out2 = torch.empty_like(fw)
    mask2 = torch.empty((b, 1, h, w), dtype=torch.float32, device=fw.device)
    blend = torch.empty((b, 3, h, w), dtype=torch.uint8, device=fw.device)
    _lib.check(lib.sb_tps_mix_blend(_lib.ptr(fw), _lib.ptr(tw), _lib.ptr(tm), _lib.ptr(o1), _lib.ptr(m1),
                                    _lib.ptr(out2), _lib.ptr(mask2), _lib.ptr(blend), b, h, w,
                                    _lib.stream_ptr()), "sb_tps_mix_blend")
    return out2, mask2, blend


def tps_warp_mask(tps_mask_channels):
    """tps_pipline.py:139-148: mean over channels, >= 0.5, then an 11x11 open of the
    INVERSE mask with cv2 border semantics, inverted back."""
    m = (tps_mask_channels.mean(dim=1, keepdim=True) >= 0.5).float()
    inv = morph_open(1.0 - m, (11, 11), border_is_zero=False)
    return 1.0 - inv


def overlap_mask(final_warp_output):
    """flowHomoAdpater.py:171-174: ``where(mean_c(mask) < 0.9, 1, 0)`` -> ``[B,H,W]``."""
    lib = _lib.load()
    fw = _lib.dev_f32(final_warp_output, "final_warp_output")
    if fw.dim() != 4 or fw.shape[1] != 6:
        raise ValueError("overlap_mask: expected [B,6,H,W]")
    b, _, h, w = fw.shape
    out = torch.empty((b, h, w), dtype=torch.float32, device=fw.device)
    _lib.check(lib.sb_overlap_mask(_lib.ptr(fw), _lib.ptr(out), b, h, w, _lib.stream_ptr()), "sb_overlap_mask")
    return out
