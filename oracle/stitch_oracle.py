"""numpy front-end of the CPU oracle (oracle.c) — TEST INFRASTRUCTURE ONLY.

A CPU restatement of the reference's hot-path arithmetic used (a) by tests/ as
the checker for the CUDA kernels, (b) by __graft_entry__.smoke(), and (c) by
bench.py's cpu_baseline / --impl reference legs as the timed CPU port.  The
product package (stitch_b200) never imports this module.

Parity status: pinned — every function here is checked against golden vectors
produced by running the reference's own Python functions
(tests/golden/make_golden.py -> tests/golden/*.npz; tests/test_oracle_golden.py).
Exceptions, which have no reference implementation to pin against:
the avg-pool pyramid C2 and the pyramid lookup C3p ("parity unpinned by the
reference": pinned only against torch's avg_pool2d / the reference's
bilinear_sampler applied with the dead code's convention).

All arrays are C-contiguous numpy float32 unless noted; names, argument meaning
and output shapes follow the reference functions cited in oracle.c.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_float, c_int, c_longlong, c_void_p

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


def _load():
    global _lib
    if _lib is None:
        import importlib.util
        spec = importlib.util.spec_from_file_location("_oracle_build", os.path.join(_HERE, "build.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        _lib = ctypes.CDLL(mod.build())
    return _lib


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a):
    return c_void_p(0) if a is None else c_void_p(a.ctypes.data)


def linspace_table(n):
    """torch.linspace(-1, 1, n) — taken from torch itself (the reference builds it with
    torch on the CPU, torch_homo_transform.py:96-99; it is an INPUT of the kernels)."""
    import torch
    return torch.linspace(-1.0, 1.0, int(n)).numpy().copy()


# ----------------------------------------------------------------- C1 / C2
def corr(fmap1, fmap2, heads=1, use_blas=True):
    """MemoryEncoder.corr (encoder.py:359-369): fp32 einsum 'bhid,bhjd->bhij'."""
    f1, f2 = _f32(fmap1), _f32(fmap2)
    b, dim, h1, w1 = f1.shape
    _, _, h2, w2 = f2.shape
    d = dim // heads
    a = f1.reshape(b * heads, d, h1 * w1)
    c = f2.reshape(b * heads, d, h2 * w2)
    if use_blas:
        vol = np.matmul(a.transpose(0, 2, 1), c)
    else:
        vol = np.empty((b * heads, h1 * w1, h2 * w2), np.float32)
        _load().o_corr(_p(a), _p(c), _p(vol), c_int(b * heads), c_int(d), c_int(h1 * w1), c_int(h2 * w2))
    return np.ascontiguousarray(vol, dtype=np.float32).reshape(b, heads, h1, w1, h2, w2)


def corr_bf16_inputs(fmap1, fmap2):
    """The same contraction on bf16-rounded operands accumulated in fp64: what the
    tcgen05 kernel computes up to fp32 accumulation order (used to separate the
    bf16 rounding contract from kernel bugs)."""
    import torch
    r1 = torch.from_numpy(_f32(fmap1)).bfloat16().double().numpy()
    r2 = torch.from_numpy(_f32(fmap2)).bfloat16().double().numpy()
    b, dim, h1, w1 = r1.shape
    _, _, h2, w2 = r2.shape
    vol = np.matmul(r1.reshape(b, dim, -1).transpose(0, 2, 1), r2.reshape(b, dim, -1))
    return vol.reshape(b, 1, h1, w1, h2, w2)


def avg_pool2x2(x):
    """F.avg_pool2d(x, 2, stride=2) on [..., H, W]."""
    x = _f32(x)
    h, w = x.shape[-2:]
    planes = int(np.prod(x.shape[:-2])) if x.ndim > 2 else 1
    out = np.empty(x.shape[:-2] + (h // 2, w // 2), np.float32)
    _load().o_avg_pool2x2(_p(x), _p(out), c_longlong(planes), c_int(h), c_int(w))
    return out


def corr_pyramid(fmap1, fmap2, num_levels=4):
    vol = corr(fmap1, fmap2)
    b, _, h1, w1, h2, w2 = vol.shape
    lv = [vol.reshape(b * h1 * w1, 1, h2, w2)]
    for _ in range(num_levels - 1):
        lv.append(avg_pool2x2(lv[-1]))
    return lv


# ----------------------------------------------------------------- C3 / C3p
def encode_flow_token(cost_maps, coords, r=4, coord_scale=1.0):
    """decoder.py:242-260 -> logical [B,(2r+1)^2,H1,W1] (returned as that view of a
    [B,H1,W1,(2r+1)^2] array, like the reference)."""
    cm, co = _f32(cost_maps), _f32(coords)
    b, _, h1, w1 = co.shape
    nq, heads, h2, w2 = cm.shape
    assert heads == 1 and nq == b * h1 * w1
    k = (2 * r + 1) ** 2
    out = np.empty((b, h1, w1, k), np.float32)
    _load().o_corr_lookup(_p(cm), _p(co), _p(out), c_int(b), c_int(h1), c_int(w1), c_int(h2), c_int(w2),
                          c_int(r), c_float(coord_scale), c_int(k), c_int(0))
    return out.transpose(0, 3, 1, 2)


def encode_flow_token_pyramid(pyramid, coords, r=4):
    outs = [encode_flow_token(cm, coords, r, 1.0 / (1 << l)) for l, cm in enumerate(pyramid)]
    return np.concatenate(outs, axis=1)


def bilinear_sampler(img, coords):
    im, co = _f32(img), _f32(coords)
    n, c, h, w = im.shape
    ho, wo = co.shape[1:3]
    out = np.empty((n, c, ho, wo), np.float32)
    _load().o_bilinear_sampler(_p(im), _p(co), _p(out), c_int(n), c_int(c), c_int(h), c_int(w), c_int(ho), c_int(wo))
    return out


# ----------------------------------------------------------------- W1
def warp(x, flo, mul_mask=None, return_overlap=False, mode="bilinear"):
    x, flo = _f32(x), _f32(flo)
    b, c, h, w = x.shape
    if mode == "nearest":
        assert mul_mask is None and not return_overlap
        out = np.empty_like(x)
        _load().o_flow_warp_nearest(_p(x), _p(flo), _p(out), c_int(b), c_int(c), c_int(h), c_int(w))
        return out
    mm = None if mul_mask is None else _f32(mul_mask)
    out = np.empty_like(x)
    ov = np.empty((b, h, w), np.float32) if return_overlap else None
    _load().o_flow_warp(_p(x), _p(flo), _p(mm), _p(out), _p(ov), c_int(b), c_int(c), c_int(h), c_int(w))
    return (out, ov) if return_overlap else out


# ----------------------------------------------------------------- W2 / W3
def homo_transformer(U, theta, out_size, return_indices=False):
    U = _f32(U)
    th = _f32(theta).reshape(-1, 3, 3)
    b, c, h, w = U.shape
    ho, wo = int(out_size[0]), int(out_size[1])
    xs, ys = linspace_table(wo), linspace_table(ho)
    out = np.empty((b, c, ho, wo), np.float32)
    idx = np.empty((b, 4, ho, wo), np.int32) if return_indices else None
    _load().o_homo_warp(_p(U), _p(th), _p(xs), _p(ys), _p(out), _p(idx), c_int(b), c_int(c), c_int(h), c_int(w),
                        c_int(ho), c_int(wo), c_int(th.shape[0]))
    return (out, idx) if return_indices else out


def tps_solve_system(source, target):
    """torch_tps_transform.py:149-185 in numpy (fp32 kernel matrix, fp64 inverse)."""
    src = _f32(source)
    tgt = _f32(target)
    b, pn, _ = src.shape
    ones = np.ones((b, pn, 1), np.float32)
    p = np.concatenate([ones, src], 2)
    diff = p.reshape(b, pn, 1, 3) - p.reshape(b, 1, pn, 3)
    d2 = np.sum(np.square(diff), 3, dtype=np.float32)
    r = d2 * np.log(d2 + np.float32(1e-6))
    w0 = np.concatenate((p, r), 2)
    w1 = np.concatenate((np.zeros((b, 3, 3), np.float32), p.transpose(0, 2, 1)), 2)
    wmat = np.concatenate((w0, w1), 1).astype(np.float64)
    w_inv = np.linalg.inv(wmat)
    tp = np.concatenate((tgt, np.zeros((b, 3, 2), np.float32)), 1).astype(np.float64)
    T = np.matmul(w_inv, tp).transpose(0, 2, 1)
    return np.ascontiguousarray(T, dtype=np.float32)


def tps_transformer(U, source, target, out_size, return_indices=False, return_coords=False, T=None):
    U = _f32(U)
    src = _f32(source)
    b, c, h, w = U.shape
    pn = src.shape[1]
    T = tps_solve_system(source, target) if T is None else _f32(T)
    ho, wo = int(out_size[0]), int(out_size[1])
    xs, ys = linspace_table(wo), linspace_table(ho)
    out = np.empty((b, c, ho, wo), np.float32)
    idx = np.empty((b, 4, ho, wo), np.int32) if return_indices else None
    crd = np.empty((b, 2, ho, wo), np.float32) if return_coords else None
    _load().o_tps_warp(_p(U), _p(T), _p(src), _p(xs), _p(ys), _p(out), _p(idx), _p(crd), c_int(b), c_int(c),
                       c_int(h), c_int(w), c_int(ho), c_int(wo), c_int(pn))
    res = [out]
    if return_indices:
        res.append(idx)
    if return_coords:
        res.append(crd)
    return res[0] if len(res) == 1 else tuple(res)


# ----------------------------------------------------------------- N1
def gma_attention(fmap, to_qk_weight, heads=1, scale=None, bf16_inputs=False):
    """Attention.forward (gma.py:54-76) in numpy: fp32 1x1 conv, fp64 contraction, softmax over keys.
    bf16_inputs=True rounds scale*q and k to bf16 first (what the tensor-core kernel contracts)."""
    import torch
    fm = _f32(fmap)
    wq = _f32(to_qk_weight).reshape(to_qk_weight.shape[0], -1)
    b, c, h, w = fm.shape
    qk = np.einsum("oc,bcn->bon", wq, fm.reshape(b, c, -1)).astype(np.float32)
    inner = qk.shape[1] // 2
    d = inner // heads
    scale = d ** -0.5 if scale is None else scale
    q = (np.float32(scale) * qk[:, :inner]).reshape(b * heads, d, -1)
    k = qk[:, inner:].reshape(b * heads, d, -1)
    n = h * w
    return gma_attention_from_qk(q, k, bf16_inputs).reshape(b, heads, n, n)


def gma_attention_from_qk(q, k, bf16_inputs=False):
    """softmax over keys of q^T k for q, k [BH, d, ...]; fp64 contraction."""
    import torch
    q, k = _f32(q), _f32(k)
    q = q.reshape(q.shape[0], q.shape[1], -1)
    k = k.reshape(k.shape[0], k.shape[1], -1)
    if bf16_inputs:
        q = torch.from_numpy(q).bfloat16().float().numpy()
        k = torch.from_numpy(k).bfloat16().float().numpy()
    sim = np.matmul(q.transpose(0, 2, 1).astype(np.float64), k.astype(np.float64))
    sim -= sim.max(axis=-1, keepdims=True)
    e = np.exp(sim)
    attn = e / e.sum(axis=-1, keepdims=True)
    return attn.astype(np.float32)


def gma_aggregate(attn, fmap, to_v_weight, gamma, heads=1):
    """Aggregate.forward (gma.py:102-115), project == None: fmap + gamma * (attn @ v)."""
    fm = _f32(fmap)
    wv = _f32(to_v_weight).reshape(to_v_weight.shape[0], -1)
    b, c, h, w = fm.shape
    n = h * w
    v = np.einsum("oc,bcn->bon", wv, fm.reshape(b, c, n)).astype(np.float32)
    d = v.shape[1] // heads
    a = _f32(attn).reshape(b * heads, n, n).astype(np.float64)
    out = np.matmul(a, v.reshape(b * heads, d, n).transpose(0, 2, 1).astype(np.float64))   # [bh, i, d]
    out = out.transpose(0, 2, 1).reshape(b, heads * d, h, w).astype(np.float32)
    return (fm + np.float32(gamma) * out).astype(np.float32)


# ----------------------------------------------------------------- N4
def _bf16(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).bfloat16().float().numpy()


def conv6x6_s2(x, w, b):
    """nn.Conv2d(kernel_size=6, stride=2, padding=2) in numpy (fp64 accumulation): x [N,C,H,W], w [O,C,6,6], b [O]."""
    x = np.asarray(x, np.float64)
    xp = np.pad(x, ((0, 0), (0, 0), (2, 2), (2, 2)))
    win = np.lib.stride_tricks.sliding_window_view(xp, (6, 6), axis=(2, 3))[:, :, ::2, ::2]     # [N,C,OH,OW,6,6]
    out = np.einsum("nchwyx,ocyx->nohw", win, np.asarray(w, np.float64), optimize=True)
    return out + np.asarray(b, np.float64).reshape(1, -1, 1, 1)


def patch_embed_proj(x, w1, b1, w2, b2, w3, b3, bf16_operands=False):
    """The conv stack of PatchEmbed.forward (encoder.py:36-43,68-73): conv -> ReLU -> conv -> ReLU -> conv,
    x [N,1,H,W] -> [N,64,H/8,W/8] fp32.  bf16_operands=True rounds the input, the weights and the two
    intermediate activations to bf16 first — exactly what the tensor-core kernel contracts — so that the kernel
    itself can be checked tightly, next to the looser contract against the fp32 reference."""
    r = _bf16 if bf16_operands else (lambda a: np.asarray(a, np.float32))
    a = np.maximum(conv6x6_s2(r(x), r(w1), b1), 0.0).astype(np.float32)
    a = np.maximum(conv6x6_s2(r(a), r(w2), b2), 0.0).astype(np.float32)
    return conv6x6_s2(r(a), r(w3), b3).astype(np.float32)


# ----------------------------------------------------------------- N3
def ccl(feature_1, feature_2, softmax_scale=10.0):
    """UDIS2Network.CCL (core/UDIS2/Homography/network.py:147-199) in numpy/fp64:
    match[q, p] = sum over c and the 3x3 offsets d of nf1[c, p+d] * nf2[c, q+d] (zero outside either
    map) — the conv2d of :161 written as nine shifted diagonals of the plain correlation."""
    f1, f2 = _f32(feature_1).astype(np.float64), _f32(feature_2).astype(np.float64)
    b, c, h, w = f1.shape
    n = h * w
    nf1 = f1 / np.maximum(np.sqrt((f1 * f1).sum(1, keepdims=True)), 1e-12)
    nf2 = f2 / np.maximum(np.sqrt((f2 * f2).sum(1, keepdims=True)), 1e-12)
    out = np.empty((b, 2, h, w), np.float32)
    ys, xs = np.divmod(np.arange(n), w)
    for i in range(b):
        c0 = nf1[i].reshape(c, n).T @ nf2[i].reshape(c, n)                    # [p, q]
        c0p = np.zeros((h + 2, w + 2, h + 2, w + 2))
        c0p[1:-1, 1:-1, 1:-1, 1:-1] = c0.reshape(h, w, h, w)
        match = np.zeros((h, w, h, w))
        for dy in (0, 1, 2):
            for dx in (0, 1, 2):
                match += c0p[dy:dy + h, dx:dx + w, dy:dy + h, dx:dx + w]
        m = match.reshape(n, n) * softmax_scale                                   # [p, q]
        m -= m.max(axis=1, keepdims=True)
        e = np.exp(m)
        prob = e / e.sum(axis=1, keepdims=True)
        out[i, 1] = (prob * (ys[None, :] - ys[:, None])).sum(1).reshape(h, w)    # flow_h
        out[i, 0] = (prob * (xs[None, :] - xs[:, None])).sum(1).reshape(h, w)    # flow_w
    return out


# ----------------------------------------------------------------- N2
def upsample_flow(flow, mask):
    """MemoryDecoder.upsample_flow (decoder.py:214-225)."""
    fl, mk = _f32(flow), _f32(mask)
    n, _, h, w = fl.shape
    out = np.empty((n, 2, 8 * h, 8 * w), np.float32)
    _load().o_upsample_flow(_p(fl), _p(mk), _p(out), c_int(n), c_int(h), c_int(w))
    return out


# ----------------------------------------------------------------- W3k
def kornia_axis_table(n):
    """create_meshgrid's normalised axis: (linspace(0, n-1, n) / (n-1) - 0.5) * 2, taken from torch."""
    import torch
    t = torch.linspace(0, n - 1, n)
    return ((t / (n - 1) - 0.5) * 2).numpy().copy()


def grid_sample(img, grid, align_corners=False):
    """F.grid_sample(img, grid, 'bilinear', 'zeros', align_corners) (kornia_tps.py:172)."""
    im, gr = _f32(img), _f32(grid)
    n, c, h, w = im.shape
    ho, wo = gr.shape[1], gr.shape[2]
    out = np.empty((n, c, ho, wo), np.float32)
    _load().o_grid_sample(_p(im), _p(gr), _p(out), c_int(n), c_int(c), c_int(h), c_int(w), c_int(ho), c_int(wo),
                          c_int(1 if align_corners else 0))
    return out


def tps_kornia_grid(kernel_centers, kernel_weights, affine_weights, h, w):
    """warp_points_tps on create_meshgrid(h, w): the sampling grid [B,H,W,2] of warp_image_tps."""
    kc, kw, aw = _f32(kernel_centers), _f32(kernel_weights), _f32(affine_weights)
    b, k, _ = kc.shape
    xs, ys = kornia_axis_table(w), kornia_axis_table(h)
    grid = np.empty((b, h, w, 2), np.float32)
    _load().o_tps_kornia_grid(_p(kc), _p(kw), _p(aw), _p(xs), _p(ys), _p(grid), c_int(b), c_int(h), c_int(w), c_int(k))
    return grid


def warp_image_tps(image, kernel_centers, kernel_weights, affine_weights, align_corners=False, return_grid=False):
    """kornia_tps.py:105-176."""
    im = _f32(image)
    grid = tps_kornia_grid(kernel_centers, kernel_weights, affine_weights, im.shape[2], im.shape[3])
    out = grid_sample(im, grid, align_corners)
    return (out, grid) if return_grid else out


# ----------------------------------------------------------------- W4 / W5
def _range(flow, mode):
    fl = _f32(flow)
    b, _, h, w = fl.shape
    out = np.empty((b, 1, h, w), np.float32)
    _load().o_range_map(_p(fl), _p(out), c_int(b), c_int(h), c_int(w), c_int(mode))
    return out


def compute_range_map(flow):
    return _range(flow, 0)


def compute_occlusion_wang(flow_ji, occlusion_are_zeros=True, threshold=False):
    if occlusion_are_zeros:
        return _range(flow_ji, 3 if threshold else 1)
    occ = _range(flow_ji, 2)
    return (occ >= 0.5).astype(np.float32) if threshold else occ


def morph_open(mask, kernel_size=(19, 19), border_is_zero=True):
    m = _f32(mask)
    h, w = m.shape[-2:]
    planes = m.size // (h * w)
    out = np.empty_like(m)
    _load().o_morph_open(_p(m), _p(out), c_int(planes), c_int(h), c_int(w), c_int(kernel_size[0]),
                         c_int(kernel_size[1]), c_int(1 if border_is_zero else 0))
    return out


def preprocess_occlusion_mask(mask, kernel_size=(19, 19)):
    return morph_open(mask, kernel_size, True)


# ----------------------------------------------------------------- W6 / W7 / W8
def composite_test_out(homo_output, homo_output2, final_warp_in, occlusion_mask=None):
    h1, h2, fw = _f32(homo_output), _f32(homo_output2), _f32(final_warp_in)
    b, _, h, w = h1.shape
    occ = None if occlusion_mask is None else _f32(occlusion_mask)
    final_warp = np.empty_like(fw)
    out2 = np.empty((b, 3, h, w), np.float32)
    m1, m2 = np.empty_like(out2), np.empty_like(out2)
    blend = np.empty((b, 3, h, w), np.uint8)
    _load().o_composite_test_out(_p(h1), _p(h2), _p(fw), _p(occ), _p(final_warp), _p(out2), _p(m1), _p(m2),
                                 _p(blend), c_int(b), c_int(h), c_int(w))
    return dict(final_warp_output=final_warp, output1=h1[:, 0:3], output2=out2, mask1=m1, mask2=m2, blend_image=blend)


def build_model_arith(warp1, warp2, mask1, mask2, net_out):
    w1, w2, m1, m2, o = map(_f32, (warp1, warp2, mask1, mask2, net_out))
    b, _, h, w = w1.shape
    lm1, lm2, st = np.empty_like(w1), np.empty_like(w1), np.empty_like(w1)
    _load().o_build_model(_p(w1), _p(w2), _p(m1), _p(m2), _p(o), _p(lm1), _p(lm2), _p(st), c_int(b), c_int(h), c_int(w))
    return dict(learned_mask1=lm1, learned_mask2=lm2, stitched_image=st)


def tps_mix_blend(final_warp, tps_warp, tps_mask, output1, mask1):
    fw, tw, tm, o1, m1 = map(_f32, (final_warp, tps_warp, tps_mask, output1, mask1))
    b, _, h, w = fw.shape
    out2 = np.empty_like(fw)
    mask2 = np.empty((b, 1, h, w), np.float32)
    blend = np.empty((b, 3, h, w), np.uint8)
    _load().o_tps_mix_blend(_p(fw), _p(tw), _p(tm), _p(o1), _p(m1), _p(out2), _p(mask2), _p(blend), c_int(b),
                            c_int(h), c_int(w))
    return out2, mask2, blend


def overlap_mask(final_warp):
    fw = _f32(final_warp)
    b, _, h, w = fw.shape
    out = np.empty((b, h, w), np.float32)
    _load().o_overlap_mask(_p(fw), _p(out), c_int(b), c_int(h), c_int(w))
    return out


# ----------------------------------------------------------------- G1 helpers (numpy)
def tensor_DLT(src_p, dst_p):
    """torch_DLT.py:17-45 in numpy fp32."""
    src, dst = _f32(src_p), _f32(dst_p)
    bs = src.shape[0]
    ones = np.ones((bs, 4, 1), np.float32)
    xy1 = np.concatenate((src, ones), 2)
    zeros = np.zeros_like(xy1)
    m1 = np.concatenate((np.concatenate((xy1, zeros), 2), np.concatenate((zeros, xy1), 2)), 2).reshape(bs, -1, 6)
    m2 = np.matmul(dst.reshape(-1, 2, 1), src.reshape(-1, 1, 2)).reshape(bs, -1, 2)
    a = np.concatenate((m1, -m2), 2)
    h8 = np.matmul(np.linalg.inv(a), dst.reshape(bs, -1, 1)).reshape(bs, 8)
    return np.concatenate((h8, ones[:, 0, :]), 1).reshape(bs, 3, 3).astype(np.float32)
