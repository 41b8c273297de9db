"""Seeded inputs of every golden case.  Shared by make_golden.py (which feeds them
to the REFERENCE's own functions here, where /root/reference exists) and by the
tests (which feed them to the oracle and to the CUDA kernels).  Everything is
drawn from torch's CPU generator, which is reproducible for a given torch build;
each .npz also stores a checksum of its inputs so drift is detected, not hidden.
"""
from __future__ import annotations

import numpy as np
import torch


def _g(seed):
    return torch.Generator().manual_seed(1234 + seed)


def checksum(*tensors) -> float:
    return float(sum(t.double().abs().sum().item() for t in tensors if isinstance(t, torch.Tensor)))


def coords_grid(b, h, w):
    ys, xs = torch.meshgrid(torch.arange(h), torch.arange(w), indexing="ij")
    return torch.stack([xs, ys], 0).float()[None].repeat(b, 1, 1, 1)


# ------------------------------------------------------------------ C1
def corr_small():
    g = _g(1)
    return dict(fmap1=torch.randn(2, 256, 6, 10, generator=g), fmap2=torch.randn(2, 256, 8, 8, generator=g))


def corr_c64():
    g = _g(2)
    return dict(fmap1=torch.randn(1, 64, 16, 16, generator=g), fmap2=torch.randn(1, 64, 16, 16, generator=g))


def corr_512():
    g = _g(3)
    return dict(fmap1=torch.randn(1, 256, 64, 64, generator=g), fmap2=torch.randn(1, 256, 64, 64, generator=g))


CORR_512_ROWS = slice(None, None, 97)
CORR_512_COLS = slice(None, None, 89)


# ------------------------------------------------------------------ C3
def lookup_small():
    g = _g(10)
    b, h1, w1, h2, w2 = 2, 3, 5, 16, 24
    cm = torch.randn(b * h1 * w1, 1, h2, w2, generator=g)
    co = torch.rand(b, 2, h1, w1, generator=g) * torch.tensor([w2 + 10.0, h2 + 10.0]).view(1, 2, 1, 1) - 5.0
    co[0, :, 0, 0] = torch.tensor([3.0, 7.0])          # exact integers (iteration-0 case)
    co[0, :, 0, 1] = torch.tensor([0.0, 0.0])
    co[0, :, 0, 2] = torch.tensor([w2 - 1.0, h2 - 1.0])
    co[1, :, 2, 4] = torch.tensor([-30.0, 50.0])       # far outside
    return dict(cost_maps=cm, coords=co)


def lookup_64():
    g = _g(11)
    b, h1, w1, h2, w2 = 1, 8, 8, 64, 64
    cm = torch.randn(b * h1 * w1, 1, h2, w2, generator=g)
    co = coords_grid(b, h1, w1) * 8.0 + torch.randn(b, 2, h1, w1, generator=g) * 2.0
    co[0, :, 0, :4] = coords_grid(1, 1, 4)[0, :, 0, :] * 8.0  # exact integers
    return dict(cost_maps=cm, coords=co)


# ------------------------------------------------------------------ W1
def warp_small():
    g = _g(20)
    x = torch.rand(2, 6, 48, 64, generator=g) * 255.0
    flo = torch.randn(2, 2, 48, 64, generator=g) * 4.0
    flo[0, :, 0, 0] = torch.tensor([-100.0, 3.0])
    flo[0, :, 5, 5] = torch.tensor([0.0, 0.0])
    flo[1, :, 47, 63] = torch.tensor([0.5, 0.5])
    return dict(x=x, flo=flo)


def warp_flow2():
    g = _g(21)
    return dict(x=torch.randn(2, 2, 40, 56, generator=g) * 3.0, flo=torch.randn(2, 2, 40, 56, generator=g) * 2.0)


def warp_512():
    g = _g(22)
    x = torch.rand(1, 6, 512, 512, generator=g) * 255.0
    lo = torch.randn(1, 2, 64, 64, generator=g) * 2.0
    flo = torch.nn.functional.interpolate(lo, size=(512, 512), mode="bilinear", align_corners=True)
    return dict(x=x, flo=flo)


WARP_512_SAMPLE = (slice(None, None, 7), slice(None, None, 5))


# ------------------------------------------------------------------ W2
def _theta_from_offsets(g, b, w, h, sigma):
    """H_mat exactly like flowHomoAdpater.py:92-108 builds it (for M at 1/8 scale)."""
    src = torch.tensor([[0.0, 0.0], [w, 0.0], [0.0, h], [w, h]]).unsqueeze(0).expand(b, -1, -1)
    dst = src + torch.randn(b, 4, 2, generator=g) * sigma
    return src, dst


def homo_small():
    g = _g(30)
    U = torch.rand(2, 6, 40, 56, generator=g) * 255.0
    U[:, 3:] = 1.0
    theta = torch.eye(3).repeat(2, 1, 1) + 0.08 * torch.randn(2, 3, 3, generator=g)
    return dict(U=U, theta=theta, out_size=(44, 60))


def homo_theta1():
    g = _g(31)
    U = (torch.rand(2, 1, 33, 47, generator=g) > 0.3).float()
    theta = (torch.eye(3) + 0.05 * torch.randn(3, 3, generator=g)).unsqueeze(0)
    return dict(U=U, theta=theta, out_size=(40, 50))


def homo_512():
    g = _g(32)
    U = torch.rand(1, 6, 512, 512, generator=g) * 255.0
    U[:, 3:] = 1.0
    theta = torch.eye(3).repeat(1, 1, 1) + 0.03 * torch.randn(1, 3, 3, generator=g)
    return dict(U=U, theta=theta, out_size=(512, 512))


HOMO_512_SAMPLE = (slice(None, None, 7), slice(None, None, 5))


def homo_degenerate():
    """t ~ 0 (point at infinity inside the canvas) and wildly out-of-range samples."""
    g = _g(33)
    U = torch.rand(1, 3, 16, 16, generator=g) * 10.0
    theta = torch.tensor([[[1.0, 0.2, 0.0], [0.1, 1.0, 0.0], [1.0, 0.0, 0.0]]])  # t = x: zero at x = 0
    return dict(U=U, theta=theta, out_size=(17, 17))


# ------------------------------------------------------------------ W3
def tps_small():
    g = _g(40)
    b, gh, gw = 1, 5, 5
    ys, xs = torch.meshgrid(torch.linspace(-1, 1, gh), torch.linspace(-1, 1, gw), indexing="ij")
    src = torch.stack([xs, ys], -1).reshape(1, -1, 2).repeat(b, 1, 1)
    tgt = src + 0.03 * torch.randn(b, gh * gw, 2, generator=g)
    # smooth image: the TPS dot product's summation order is unspecified (BLAS), so
    # coordinates agree only to ~1e-5; a smooth image keeps the 1e-3 contract meaningful
    yy, xx = torch.meshgrid(torch.linspace(0, 3.0, 32), torch.linspace(0, 4.0, 40), indexing="ij")
    base = torch.stack([torch.sin(xx) + yy, torch.cos(yy * 1.3) * 2, xx * yy * 0.3, torch.ones_like(xx),
                        torch.ones_like(xx), torch.ones_like(xx)], 0)
    return dict(U=base[None].repeat(b, 1, 1, 1).contiguous(), source=src, target=tgt, out_size=(32, 40))


# ------------------------------------------------------------------ N1
def gma_small():
    """Attention / Aggregate as decoder.py:197 / gru.py:316 build them: dim = 128, heads = 1, dim_head = 128."""
    g = _g(70)
    fmap = torch.randn(2, 128, 12, 16, generator=g)
    motion = torch.randn(2, 128, 12, 16, generator=g)
    w_qk = (torch.rand(256, 128, 1, 1, generator=g) * 2 - 1) * (3.0 / 128 ** 0.5)   # sharper than the default init
    w_v = (torch.rand(128, 128, 1, 1, generator=g) * 2 - 1) * (1.0 / 128 ** 0.5)
    return dict(fmap=fmap, motion=motion, w_qk=w_qk, w_v=w_v, gamma=torch.tensor([0.7]))


# ------------------------------------------------------------------ N3
def ccl_small():
    """Post-ReLU-like (non-negative) features with a shifted copy so the match volume has structure."""
    g = _g(80)
    f1 = torch.relu(torch.randn(2, 64, 10, 12, generator=g))
    f2 = torch.roll(f1, shifts=(1, -2), dims=(2, 3)) + 0.3 * torch.relu(torch.randn(2, 64, 10, 12, generator=g))
    return dict(feature_1=f1, feature_2=f2)


# ------------------------------------------------------------------ N2
def upsample_small():
    g = _g(60)
    return dict(flow=torch.randn(2, 2, 12, 20, generator=g) * 3.0, mask=torch.randn(2, 576, 12, 20, generator=g) * 2.0)


# ------------------------------------------------------------------ N4
def patch_embed_small():
    """Six 64x64 cost maps with the statistics of a real volume (dot products of 256 N(0,1) channels: std 16) and
    nn.Conv2d-style seeded weights U(-1/sqrt(fan_in), 1/sqrt(fan_in))."""
    g = _g(60)
    x = torch.randn(6, 1, 64, 64, generator=g) * 16.0
    x[5] = 0.0                                             # an all-zero map: output = the bias chain alone

    def conv(o, c):
        bound = 1.0 / (c * 36) ** 0.5
        return ((torch.rand(o, c, 6, 6, generator=g) * 2 - 1) * bound, (torch.rand(o, generator=g) * 2 - 1) * bound)
    (w1, b1), (w2, b2), (w3, b3) = conv(16, 1), conv(32, 16), conv(64, 32)
    return dict(x=x, w1=w1, b1=b1, w2=w2, b2=b2, w3=w3, b3=b3)


# ------------------------------------------------------------------ W3k
def tps_kornia_small():
    """warp_image_tps inputs as tps_pipline.py:364-381 builds them: control points in pixels
    divided by the canvas size (so in [0,1]), reverse transform dst -> src."""
    g = _g(45)
    b, gh, gw, h, w = 2, 5, 5, 36, 44
    ys, xs = torch.meshgrid(torch.linspace(0.05, 0.95, gh), torch.linspace(0.05, 0.95, gw), indexing="ij")
    src = torch.stack([xs, ys], -1).reshape(1, -1, 2).repeat(b, 1, 1)
    dst = src + 0.02 * torch.randn(b, gh * gw, 2, generator=g)
    yy, xx = torch.meshgrid(torch.linspace(0, 3.0, h), torch.linspace(0, 4.0, w), indexing="ij")
    base = torch.stack([torch.sin(xx) + yy, torch.cos(yy * 1.3) * 2, xx * yy * 0.3, torch.ones_like(xx),
                        torch.ones_like(xx), torch.ones_like(xx)], 0)
    return dict(image=base[None].repeat(b, 1, 1, 1).contiguous(), points_src=src, points_dst=dst)


# ------------------------------------------------------------------ W4 / W5
def range_small():
    g = _g(50)
    f_ij = torch.randn(2, 2, 48, 64, generator=g) * 3.0
    f_ji = torch.randn(2, 2, 48, 64, generator=g) * 3.0
    f_ji[0, :, :8, :8] = 0.0          # exact-integer targets
    f_ji[1, :, 10:20, 10:20] = 500.0  # far out of the image
    return dict(flow_ij=f_ij, flow_ji=f_ji)


def range_smooth():
    g = _g(51)
    lo = torch.randn(2, 2, 2, 8, 8, generator=g) * 2.0
    up = [torch.nn.functional.interpolate(lo[i], size=(64, 64), mode="bilinear", align_corners=True) for i in range(2)]
    return dict(flow_ij=up[0], flow_ji=up[1])


def morph_small():
    g = _g(60)
    lo = torch.rand(2, 1, 8, 10, generator=g)
    m = torch.nn.functional.interpolate(lo, size=(64, 80), mode="bilinear", align_corners=False)
    m[1, 0, :30, :] = 1.0
    return dict(mask=m * 1.2)


def morph_big():
    g = _g(61)
    lo = torch.rand(1, 1, 12, 20, generator=g)
    m = torch.nn.functional.interpolate(lo, size=(211, 397), mode="bilinear", align_corners=False)
    return dict(mask=(m > 0.35).float())


# ------------------------------------------------------------------ W7
def build_model_small():
    g = _g(70)
    s = (1, 3, 24, 40)
    return dict(warp1=torch.rand(s, generator=g) * 2 - 1, warp2=torch.rand(s, generator=g) * 2 - 1,
                mask1=(torch.rand(s, generator=g) > 0.3).float(), mask2=(torch.rand(s, generator=g) > 0.3).float(),
                net_out=torch.rand(1, 1, 24, 40, generator=g))


# ------------------------------------------------------------------ W8
def tps_mix_small():
    g = _g(80)
    s = (1, 3, 40, 48)
    m1 = torch.zeros(s)
    m1[..., :, :30] = 1.0
    fw = torch.rand(s, generator=g) * 255.0
    fw[..., 20:, :] = 0.0
    tmask3 = torch.zeros(1, 3, 40, 48)
    tmask3[..., 4:38, 10:46] = 1.0
    tmask3[..., 15:17, 20:22] = 0.0   # small hole the 11x11 open of the inverse removes
    return dict(final_warp=fw, tps_warp_raw=torch.rand(s, generator=g) * 255.0, tps_mask3=tmask3,
                output1=torch.rand(s, generator=g) * 255.0, mask1=m1)


# ------------------------------------------------------------------ adapter
class StubHomo(torch.nn.Module):
    """Deterministic stand-in for UDIS2Network: fixed 4-point offsets."""

    def __init__(self, offsets):
        super().__init__()
        self.register_buffer("offsets", offsets)

    def forward(self, a, b):
        return self.offsets[: a.shape[0]].reshape(a.shape[0], -1).to(a.device), None


class StubFlow(torch.nn.Module):
    """Deterministic stand-in for FlowFormer: returns a stored flow, alternating
    between the forward and the backward field on successive calls."""

    def __init__(self, flows):
        super().__init__()
        self.flows = flows
        self.calls = 0
        self.eval()

    def forward(self, a, b, out_dict=None):
        f = self.flows[self.calls % len(self.flows)].to(a.device)
        self.calls += 1
        return [f[: a.shape[0]].clone()]


class Cfg(dict):
    __getattr__ = dict.get

    def __hasattr__(self, k):
        return k in self


class AttrCfg:
    """Stand-in for yacs.CfgNode: attribute access + hasattr semantics."""

    def __init__(self, **kw):
        self.__dict__.update(kw)


def adapter_cfg():
    return AttrCfg(use_forward=False, use_combine_h_flow=False, use_fb_consistency_mask=True,
                   test_not_use_combine_h_flow=True, only_homo=False)


def _smooth_flow(g, b, size, sigma):
    lo = torch.randn(b, 2, size // 8, size // 8, generator=g) * sigma
    return torch.nn.functional.interpolate(lo, size=(size, size), mode="bilinear", align_corners=True)


def adapter_train_eval():
    g = _g(90)
    b, s = 2, 128
    im1 = torch.rand(b, 3, s, s, generator=g) * 255.0
    im2 = torch.rand(b, 3, s, s, generator=g) * 255.0
    offsets = torch.randn(b, 4, 2, generator=g) * 6.0
    flows = [_smooth_flow(g, b, s, 2.0), _smooth_flow(g, b, s, 2.0)]
    return dict(image1=im1, image2=im2, offsets=offsets, flows=flows)


def adapter_test_out():
    g = _g(91)
    b, s = 1, 512
    # smooth images: adapter-level parity crosses a GPU-vs-CPU 3x3 inverse, so the
    # warp coordinates agree to ~1e-5 px, not bit-for-bit
    yy, xx = torch.meshgrid(torch.linspace(0, 6.0, s), torch.linspace(0, 5.0, s), indexing="ij")
    im1 = (torch.stack([torch.sin(xx) + torch.cos(yy), torch.sin(xx * yy * 0.3), torch.cos(xx - yy)], 0) * 60 + 128)[None]
    im2 = (torch.stack([torch.cos(xx * 1.1) + torch.sin(yy), torch.cos(xx * yy * 0.2), torch.sin(xx + yy)], 0) * 60 + 128)[None]
    offsets = torch.randn(b, 4, 2, generator=g) * 20.0
    flows = [_smooth_flow(g, b, s, 2.0), _smooth_flow(g, b, s, 2.0)]
    return dict(image1=im1.contiguous(), image2=im2.contiguous(), offsets=offsets, flows=flows)


def demo1_pair():
    """BASELINE config 1: the reference's demo/demo1 pair (tests/golden/demo1/input{1,2}.jpg, copies of the
    reference's fixture files), decoded as out.py:129-146 does (cv2 BGR -> RGB, float 0..255, [1,3,512,512]),
    with the SURVEY 8(d) stand-ins for the networks: 4-point offsets N(0, 20^2) px, smooth residual flows
    N(0, 2^2) at 1/8 resolution upsampled x8, a 13x13 TPS mesh with N(0, 0.02^2) targets."""
    import os
    import cv2
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "demo1")
    ims = []
    for name in ("input1.jpg", "input2.jpg"):
        im = cv2.cvtColor(cv2.imread(os.path.join(here, name)).astype("uint8"), cv2.COLOR_BGR2RGB)
        ims.append(torch.from_numpy(np.array(im).astype(np.uint8)[..., :3]).permute(2, 0, 1).float().unsqueeze(0).contiguous())
    g = _g(95)
    offsets = torch.randn(1, 4, 2, generator=g) * 20.0
    flows = [_smooth_flow(g, 1, 512, 2.0), _smooth_flow(g, 1, 512, 2.0)]
    ys, xs = torch.meshgrid(torch.linspace(-1, 1, 13), torch.linspace(-1, 1, 13), indexing="ij")
    src = torch.stack([xs, ys], -1).reshape(1, -1, 2)
    tgt = src + 0.02 * torch.randn(1, 169, 2, generator=g)
    return dict(image1=ims[0], image2=ims[1], offsets=offsets, flows=flows, tps_source=src, tps_target=tgt)


DEMO1_SAMPLE = (slice(2, None, 5), slice(1, None, 5))
ADAPTER_SAMPLE = (slice(None, None, 3), slice(None, None, 3))
# pixels of the test_out canvas whose compositing inputs AND outputs are stored (tests/golden/composite_w6.npz)
W6_SAMPLE = (slice(1, None, 4), slice(2, None, 4))


def to_numpy(d):
    out = {}
    for k, v in d.items():
        if isinstance(v, torch.Tensor):
            out[k] = v.numpy()
        elif isinstance(v, (list, tuple)) and v and isinstance(v[0], torch.Tensor):
            out[k] = [t.numpy() for t in v]
        else:
            out[k] = v
    return out
