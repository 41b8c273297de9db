"""Generate tests/golden/*.npz by running the REFERENCE's own Python functions.

Run here (the build container), where the reference is mounted read-only at
/root/reference:      python tests/golden/make_golden.py
The GPU box has no /root/reference: tests only read the committed .npz files.

Import shims: fake `skimage`, fake `timm.*` (import-time only,
none of the stubbed symbols executes on the path), `.cuda()` -> identity on this
GPU-less box, and an attribute-bag config in place of yacs.CfgNode.
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("STITCH_REFERENCE", "/root/reference")
sys.path.insert(0, HERE)
import cases  # noqa: E402


def install_shims():
    sys.path[:0] = [REF, os.path.join(REF, "core")]
    for name in ("skimage", "skimage.io"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["skimage"].io = sys.modules["skimage.io"]

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    class _Dummy(torch.nn.Module):
        def __init__(self, *a, **k):
            super().__init__()

    timm = mod("timm", create_model=lambda *a, **k: None)
    mod("timm.data", IMAGENET_DEFAULT_MEAN=(0.485, 0.456, 0.406), IMAGENET_DEFAULT_STD=(0.229, 0.224, 0.225))
    layers = mod("timm.models.layers", Mlp=_Dummy, DropPath=_Dummy, to_2tuple=lambda x: (x, x),
                 trunc_normal_=lambda *a, **k: None, activations=types.SimpleNamespace())
    models = mod("timm.models", layers=layers)
    mod("timm.models.registry", register_model=lambda f: f)
    mod("timm.models.vision_transformer", Attention=_Dummy, Block=_Dummy, _cfg=lambda **k: {})
    timm.models = models
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self


def capture_gather_indices(fn, *args, **kw):
    """Run fn while recording the index tensors of its four torch.gather calls
    (idx_a..idx_d of the UDIS sampler, torch_homo_transform.py:54-79)."""
    rec = []
    orig = torch.gather

    def spy(inp, dim, index, *a, **k):
        rec.append(index[:, 0].clone())
        return orig(inp, dim, index, *a, **k)

    torch.gather = spy
    try:
        out = fn(*args, **kw)
    finally:
        torch.gather = orig
    return out, rec


def decode_indices(rec, b, h, w, hout, wout):
    """idx_a = base + y0*W + x0, idx_b = base + y1*W + x0, idx_c = base + y0*W + x1 -> (x0,x1,y0,y1)."""
    ia, ib, ic, _ = [r.reshape(b, hout, wout) for r in rec]
    base = (torch.arange(b) * (h * w)).view(b, 1, 1)
    ia, ib, ic = ia - base, ib - base, ic - base
    x0, y0 = ia % w, ia // w
    y1 = ib // w
    x1 = ic % w
    return torch.stack([x0, x1, y0, y1], 1).to(torch.int32).numpy()


def save(name, inputs_checksum, **arrays):
    path = os.path.join(os.environ.get("STITCH_GOLDEN_OUT", HERE), name + ".npz")
    np.savez_compressed(path, inputs_checksum=np.float64(inputs_checksum), **arrays)
    print(f"{name}: {os.path.getsize(path) / 1024:.1f} KiB")


def main():
    install_shims()
    from types import SimpleNamespace
    import core.warp_utils as ref_wu
    import core.udis_utils.torch_homo_transform as ref_homo
    import core.udis_utils.torch_tps_transform as ref_tps
    import core.udis_utils.torch_DLT as ref_dlt
    import core.utils.utils as ref_utils
    from core.FlowFormer.PerCostFormer3.encoder import MemoryEncoder
    from core.FlowFormer.PerCostFormer3.decoder import MemoryDecoder
    import core.flowHomoAdpater as ref_ad
    from core.UDIS2.Composition.network import build_model as ref_build_model

    enc_self = SimpleNamespace(cfg=SimpleNamespace(cost_heads_num=1))
    torch.set_grad_enabled(False)

    # ---------------------------------------------------------------- C1
    for name, fn in (("corr_small", cases.corr_small), ("corr_c64", cases.corr_c64)):
        c = fn()
        vol = MemoryEncoder.corr(enc_self, c["fmap1"], c["fmap2"])
        save(name, cases.checksum(*c.values()), vol=vol.numpy())
    c = cases.corr_512()
    vol = MemoryEncoder.corr(enc_self, c["fmap1"], c["fmap2"])
    v2 = vol.reshape(4096, 4096)
    cm = vol.permute(0, 2, 3, 1, 4, 5).contiguous().view(4096, 1, 64, 64)      # encoder.py:260
    # pyramid (C2): no reference implementation — F.avg_pool2d chained (encoder.py:376 hint)
    l1 = F.avg_pool2d(cm, 2, stride=2); l2 = F.avg_pool2d(l1, 2, stride=2); l3 = F.avg_pool2d(l2, 2, stride=2)
    save("corr_512", cases.checksum(*c.values()), vol_sample=v2[cases.CORR_512_ROWS, cases.CORR_512_COLS].numpy(),
         vol_absmax=np.float32(v2.abs().max()), vol_rms=np.float32(v2.pow(2).mean().sqrt()),
         lvl1_sample=l1[::97].numpy(), lvl2_sample=l2[::97].numpy(), lvl3_sample=l3[::97].numpy())

    # ---------------------------------------------------------------- C3
    for name, fn in (("lookup_small", cases.lookup_small), ("lookup_64", cases.lookup_64)):
        c = fn()
        out = MemoryDecoder.encode_flow_token(None, c["cost_maps"], c["coords"])
        extra = {}
        if name == "lookup_small":
            for rr in (0, 1, 2, 7):      # radii of the non-default branches (decoder.py:295-315); r = 0: the centre only
                extra[f"out_r{rr}"] = MemoryDecoder.encode_flow_token(None, c["cost_maps"], c["coords"], r=rr).contiguous().numpy()
            # C3p convention (common.py:245-248): centroid / 2**i + delta on pooled maps
            pyr = [c["cost_maps"]]
            for _ in range(2):
                pyr.append(F.avg_pool2d(pyr[-1], 2, stride=2))
            outs = [MemoryDecoder.encode_flow_token(None, pm, c["coords"] / 2 ** i) for i, pm in enumerate(pyr)]
            extra["out_pyramid"] = torch.cat(outs, 1).numpy()
        save(name, cases.checksum(*c.values()), out=out.contiguous().numpy(),
             out_strides=np.array(out.stride()), **extra)
    c = cases.lookup_small()
    pts = torch.rand(3, 7, 9, 2, generator=torch.Generator().manual_seed(5)) * 30 - 3
    img = torch.randn(3, 4, 16, 24, generator=torch.Generator().manual_seed(6))
    samp, msk = ref_utils.bilinear_sampler(img, pts, mask=True)
    save("bilinear_sampler", cases.checksum(img, pts), img=img.numpy(), pts=pts.numpy(), out=samp.numpy(), mask=msk.numpy())

    # ---------------------------------------------------------------- W1
    for name, fn in (("warp_small", cases.warp_small), ("warp_flow2", cases.warp_flow2)):
        c = fn()
        save(name, cases.checksum(*c.values()), out=ref_wu.warp(c["x"], c["flo"]).numpy(),
             out_nearest=ref_wu.warp(c["x"], c["flo"], mode="nearest").numpy())
    c = cases.warp_512()
    out = ref_wu.warp(c["x"], c["flo"])
    save("warp_512", cases.checksum(*c.values()), out_sample=out[..., cases.WARP_512_SAMPLE[0], cases.WARP_512_SAMPLE[1]].numpy())

    # ---------------------------------------------------------------- W2
    for name, fn in (("homo_small", cases.homo_small), ("homo_theta1", cases.homo_theta1),
                     ("homo_degenerate", cases.homo_degenerate)):
        c = fn()
        out, rec = capture_gather_indices(ref_homo.transformer, c["U"], c["theta"], c["out_size"])
        b, _, h, w = c["U"].shape
        idx = decode_indices(rec, b, h, w, *c["out_size"])
        save(name, cases.checksum(c["U"], c["theta"]), out=out.contiguous().numpy(), idx=idx)
    c = cases.homo_512()
    out, rec = capture_gather_indices(ref_homo.transformer, c["U"], c["theta"], c["out_size"])
    idx = decode_indices(rec, 1, 512, 512, 512, 512)
    sy, sx = cases.HOMO_512_SAMPLE
    save("homo_512", cases.checksum(c["U"], c["theta"]), out_sample=out[..., sy, sx].contiguous().numpy(),
         idx_sample=idx[..., sy, sx], mask_bits=np.packbits((out[0, 3] > 0.5).numpy()))

    # ---------------------------------------------------------------- W3
    c = cases.tps_small()
    out, rec = capture_gather_indices(ref_tps.transformer, c["U"], c["source"], c["target"], c["out_size"])
    b, _, h, w = c["U"].shape
    idx = decode_indices(rec, b, h, w, *c["out_size"])
    save("tps_small", cases.checksum(c["U"], c["source"], c["target"]), out=out.contiguous().numpy(), idx=idx)

    # ---------------------------------------------------------------- W4
    for name, fn in (("range_small", cases.range_small), ("range_smooth", cases.range_smooth)):
        c = fn()
        rm = ref_wu.compute_range_map(c["flow_ji"])
        occ = ref_wu.compute_occlusion(c["flow_ij"], c["flow_ji"], "wang", occlusion_are_zeros=True, boundaries_occluded=True)
        occ_nz = ref_wu.compute_occlusion(c["flow_ij"], c["flow_ji"], "wang", occlusion_are_zeros=False, boundaries_occluded=True)
        occ_nb = ref_wu.compute_occlusion(c["flow_ij"], c["flow_ji"], "wang", occlusion_are_zeros=True, boundaries_occluded=False)
        brox = ref_wu.compute_occlusion(c["flow_ij"], c["flow_ji"], "brox")
        save(name, cases.checksum(*c.values()), range_map=rm.numpy(), occ=occ.numpy(), occ_nz=occ_nz.numpy(),
             occ_nb=occ_nb.numpy(), brox=brox.numpy())

    # ---------------------------------------------------------------- W5
    for name, fn in (("morph_small", cases.morph_small), ("morph_big", cases.morph_big)):
        c = fn()
        out = ref_ad.preprocess_occlusion_mask(c["mask"])
        out7 = ref_ad.preprocess_occlusion_mask(c["mask"], kernel_size=(7, 11))
        save(name, cases.checksum(c["mask"]), out_bits=np.packbits(out.numpy() > 0.5), out7_bits=np.packbits(out7.numpy() > 0.5),
             shape=np.array(out.shape))

    # ---------------------------------------------------------------- W7
    c = cases.build_model_small()
    r = ref_build_model(lambda *a: c["net_out"], c["warp1"], c["warp2"], c["mask1"], c["mask2"])
    save("build_model", cases.checksum(*c.values()), **{k: v.numpy() for k, v in r.items()})

    # ---------------------------------------------------------------- N1
    import core.FlowFormer.PerCostFormer3.gma as ref_gma
    c = cases.gma_small()
    att = ref_gma.Attention(args=None, dim=128, heads=1, max_pos_size=160, dim_head=128)
    agg = ref_gma.Aggregate(args=None, dim=128, dim_head=128, heads=1)
    att.to_qk.weight.data.copy_(c["w_qk"]); agg.to_v.weight.data.copy_(c["w_v"]); agg.gamma.data.copy_(c["gamma"])
    attn = att(c["fmap"])
    save("gma_small", cases.checksum(*c.values()), attn=attn.numpy(), out=agg(attn, c["motion"]).numpy())

    # ---------------------------------------------------------------- N3
    from core.UDIS2.Homography.network import UDIS2Network
    stub = SimpleNamespace(extract_patches=lambda x, kernel=3, stride=1: UDIS2Network.extract_patches(None, x, kernel, stride))
    c = cases.ccl_small()
    save("ccl_small", cases.checksum(*c.values()), flow=UDIS2Network.CCL(stub, c["feature_1"], c["feature_2"]).numpy())

    # ---------------------------------------------------------------- N4: the reference's own PatchEmbed (encoder.py:20-92)
    # built as MemoryEncoder builds it (:179: patch_size 8, embed_dim = cost_latent_input_dim = 64, pe 'linear',
    # patch_embed 'single'); the conv stack's output is captured with a forward hook on its last layer while
    # the module's forward runs, the module's final tokens are stored too (for the drop-in forward body).
    from core.FlowFormer.PerCostFormer3.encoder import PatchEmbed
    c = cases.patch_embed_small()
    torch.manual_seed(1234)
    pe_mod = PatchEmbed(patch_size=8, in_chans=1, embed_dim=64, pe="linear",
                        cfg=SimpleNamespace(patch_embed="single", use_rpe=False)).eval()
    for idx, k in ((0, "1"), (2, "2"), (4, "3")):
        pe_mod.proj[idx].weight.data.copy_(c["w" + k]); pe_mod.proj[idx].bias.data.copy_(c["b" + k])
    grabbed = []
    hook = pe_mod.proj[4].register_forward_hook(lambda m, i, o: grabbed.append(o.clone()))
    tokens, size = pe_mod(c["x"])
    hook.remove()
    save("patch_embed_small", cases.checksum(*c.values()), proj=grabbed[0].numpy(), tokens=tokens.numpy(),
         size=np.array(size), ffn0_w=pe_mod.ffn_with_coord[0].weight.detach().numpy(), ffn0_b=pe_mod.ffn_with_coord[0].bias.detach().numpy(),
         ffn2_w=pe_mod.ffn_with_coord[2].weight.detach().numpy(), ffn2_b=pe_mod.ffn_with_coord[2].bias.detach().numpy())

    # ---------------------------------------------------------------- N2
    c = cases.upsample_small()
    save("upsample_small", cases.checksum(*c.values()), out=MemoryDecoder.upsample_flow(None, c["flow"], c["mask"]).numpy())

    # ---------------------------------------------------------------- W3k (kornia absent: the two
    # functions the reference imports from it are restated from kornia's published source in a stub
    # module; everything else — _pair_square_euclidean, _kernel_distance, custom_get_tps_transform,
    # warp_image_tps — is the reference's own vendored code in core/inference/tps_methods/kornia_tps.py)
    def _k_warp_points_tps(points_src, kernel_centers, kernel_weights, affine_weights):
        pair = ref_ktps._pair_square_euclidean(points_src, kernel_centers)
        k_matrix = ref_ktps._kernel_distance(pair)
        return (k_matrix[..., None].mul(kernel_weights[:, None]).sum(-2)
                + points_src[..., None].mul(affine_weights[:, None, 1:]).sum(-2)
                + affine_weights[:, None, 0])

    def _k_create_meshgrid(height, width, normalized_coordinates=True, device=None, dtype=torch.float32):
        xs = torch.linspace(0, width - 1, width, device=device, dtype=dtype)
        ys = torch.linspace(0, height - 1, height, device=device, dtype=dtype)
        if normalized_coordinates:
            xs = (xs / (width - 1) - 0.5) * 2
            ys = (ys / (height - 1) - 0.5) * 2
        return torch.stack(torch.meshgrid([xs, ys], indexing="ij"), dim=-1).permute(1, 0, 2).unsqueeze(0)

    for name in ("kornia", "kornia.geometry", "kornia.geometry.transform", "kornia.utils", "kornia.core"):
        sys.modules.setdefault(name, types.ModuleType(name))
    kt = sys.modules["kornia.geometry.transform"]
    kt.warp_image_tps = kt.get_tps_transform = None          # shadowed by the reference's own definitions
    kt.warp_points_tps = _k_warp_points_tps
    sys.modules["kornia.utils"].create_meshgrid = _k_create_meshgrid
    sys.modules["kornia.core"].Tensor = torch.Tensor
    import core.inference.tps_methods.kornia_tps as ref_ktps
    c = cases.tps_kornia_small()
    kwt, awt = ref_ktps.custom_get_tps_transform(c["points_dst"], c["points_src"])
    arrays = dict(kernel_weights=kwt.numpy(), affine_weights=awt.numpy())
    for ac in (False, True):
        arrays[f"out_ac{int(ac)}"] = ref_ktps.warp_image_tps(c["image"], c["points_src"], kwt, awt, align_corners=ac).numpy()
    coords = _k_create_meshgrid(c["image"].shape[2], c["image"].shape[3]).reshape(-1, 2).expand(c["image"].shape[0], -1, -1)
    arrays["grid"] = _k_warp_points_tps(coords, c["points_src"], kwt, awt).view(-1, c["image"].shape[2], c["image"].shape[3], 2).numpy()
    save("tps_kornia", cases.checksum(*c.values()), **arrays)

    # ---------------------------------------------------------------- W8: the reference's own
    # core/inference/tps_pipline.py:tps_H_warp, lines :138-170 executed as written.  matplotlib is absent
    # (import-time stub; nothing of it runs with is_plot=False); the three stages UPSTREAM of :138
    # (preprocess / sample_init_points / warp_by_tps: point sampling + the third-party OpenCV TPS, out of
    # scope) are replaced by stubs that hand the case's seeded tensors to the real mix / open / blend code.
    for name in ("matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    import core.inference.tps_pipline as ref_tp
    c = cases.tps_mix_small()
    h, w = c["final_warp"].shape[-2:]
    saved = (ref_tp.preprocess, ref_tp.sample_init_points, ref_tp.warp_by_tps, ref_tp.cv2)
    opened = []                                  # result of the cv2.dilate at :146 (the opened inverse mask)

    def _dilate(*a, **k):
        opened.append(saved[3].dilate(*a, **k))
        return opened[-1]
    ref_tp.cv2 = SimpleNamespace(getStructuringElement=saved[3].getStructuringElement, MORPH_RECT=saved[3].MORPH_RECT,
                                 erode=saved[3].erode, dilate=_dilate)
    ref_tp.preprocess = lambda residual_flow, valid, **k: residual_flow
    ref_tp.sample_init_points = lambda residual_flow, **k: (None, None, torch.zeros(1, 4, 2), torch.zeros(1, 4, 2))
    ref_tp.warp_by_tps = lambda H_warp, H_warp_mask, ps, pd, **k: torch.cat((c["tps_warp_raw"].clone(), c["tps_mask3"].clone()), 1)
    try:
        od = ref_tp.tps_H_warp(
            SimpleNamespace(output1=c["output1"].clone(), mask1=c["mask1"].clone(), H_warp=None, H_warp_mask=None,
                            final_warp=c["final_warp"].clone(), mask2=None, residual_flow=torch.zeros(1, 2, h, w),
                            valid=None, occlusion_mask=None, border_points_mask=None),
            SimpleNamespace(width_min=0, height_min=0, out_height=h, out_width=w),
            SimpleNamespace(grid_h=12, grid_w=12, pad_num=0, residual_flow_use_forward=False, flow_limit=None,
                            add_corner=False, get_pt_methods=None, add_meshgrid=False, affine_scale=1.0,
                            kernel_scale=1.0, use_boundary_limit=False, tps_method="opencv",
                            output2_is_only_tps=False, do_avg_pooling=False),
            inpaint_fn=None, is_plot=False)
    finally:
        ref_tp.preprocess, ref_tp.sample_init_points, ref_tp.warp_by_tps, ref_tp.cv2 = saved
    assert len(opened) == 1
    tm = 1.0 - torch.tensor(opened[0])[None, None]                               # :147-148
    assert torch.equal(od["tps_output"], c["tps_warp_raw"] * tm)
    save("tps_mix", cases.checksum(*c.values()), tps_mask=tm.numpy(), tps_output=od["tps_output"].numpy(),
         output2=od["output2"].numpy(), mask2=od["mask2"].numpy(), blend=od["new_blend_image"].numpy())

    # ---------------------------------------------------------------- G1
    g = torch.Generator().manual_seed(77)
    src = torch.tensor([[0.0, 0.0], [64, 0.0], [0.0, 48], [64, 48]]).unsqueeze(0).expand(3, -1, -1)
    dst = src + torch.randn(3, 4, 2, generator=g) * 4
    Hm = ref_dlt.tensor_DLT(src, dst)
    mesh = ref_wu.H2Mesh(Hm, ref_wu.get_rigid_mesh(3, 48, 64, 7, 9), 7, 9)
    fl = torch.randn(2, 2, 16, 20, generator=g)
    save("geometry", cases.checksum(dst), dst=dst.numpy(), H=Hm.numpy(), mesh=mesh.numpy(), flow=fl.numpy(),
         flow_resized=ref_wu.resize_flow(fl.clone(), (37, 45)).numpy())

    # ---------------------------------------------------------------- adapter forwards (stub networks)
    c = cases.adapter_train_eval()
    ad = ref_ad.FlowHomoAdpater(cases.StubHomo(c["offsets"]), cases.StubFlow(c["flows"]), cases.adapter_cfg())
    ad.eval()
    od = ad.train_eval_foward(c["image1"], c["image2"])
    save("adapter_train_eval", cases.checksum(c["image1"], c["image2"], c["offsets"], *c["flows"]),
         output_H=od["output_H"].contiguous().numpy(), output_H_inv=od["output_H_inv"].contiguous().numpy(),
         final_warp_output=od["final_warp_output"].numpy(), overlap=od["overlap"].numpy(),
         origin_occlusion_mask=od["origin_occlusion_mask"].numpy(), H=od["H"].numpy())

    c = cases.adapter_test_out()
    ad = ref_ad.FlowHomoAdpater(cases.StubHomo(c["offsets"]), cases.StubFlow(c["flows"]), cases.adapter_cfg())
    ad.eval()
    # W6: record what the reference's own compositing block (flowHomoAdpater.py:317,339-360) consumes, by
    # spying on the three functions whose results feed it — transformer calls #2..#5 of test_out_forward are
    # homo_output (:292), homo_output2 (:310), residual_flow_output (:314) and occlusion_mask (:335); warp is
    # called once (:316); preprocess_occlusion_mask's second call (:336) yields the mask that multiplies (:339)
    spy = dict(transformer=[], warp=[], morph=[])
    orig = (ref_ad.torch_homo_transform.transformer, ref_ad.warp, ref_ad.preprocess_occlusion_mask)

    def _spy(key, fn):
        def wrapped(*a, **k):
            r = fn(*a, **k)
            spy[key].append(r.clone())
            return r
        return wrapped

    ref_ad.torch_homo_transform.transformer = _spy("transformer", orig[0])
    ref_ad.warp = _spy("warp", orig[1])
    ref_ad.preprocess_occlusion_mask = _spy("morph", orig[2])
    try:
        od = ad.test_out_forward(c["image1"], c["image2"])
    finally:
        ref_ad.torch_homo_transform.transformer, ref_ad.warp, ref_ad.preprocess_occlusion_mask = orig
    assert len(spy["transformer"]) == 5 and len(spy["warp"]) == 1 and len(spy["morph"]) == 2
    w6 = dict(homo_output=spy["transformer"][1], homo_output2=spy["transformer"][2],
              flow_mask=spy["transformer"][3][:, 2:3], warp_out=spy["warp"][0], occlusion_mask=spy["morph"][1])
    w6["final_warp_in"] = w6["warp_out"] * w6["flow_mask"]                       # :317, the reference's own op
    assert torch.equal(w6["occlusion_mask"], od["occlusion_mask"])
    assert torch.equal((w6["final_warp_in"] * w6["occlusion_mask"])[:, 0:3], od["final_warp"])   # :339
    wy, wx = cases.W6_SAMPLE                 # compositing is per pixel: a pixel subset is a complete test vector
    w6_arrays = {k: v[..., wy, wx].contiguous().numpy() for k, v in w6.items()}
    for k in ("final_warp", "output1", "output2", "mask1", "mask2", "blend_image"):
        w6_arrays["out_" + k] = od[k][..., wy, wx].contiguous().numpy()
    save("composite_w6", cases.checksum(c["image1"], c["image2"], c["offsets"], *c["flows"]), **w6_arrays)
    sy, sx = cases.ADAPTER_SAMPLE
    arrays = {}
    for k in ("H_warp", "final_warp", "output1", "output2", "mask1", "mask2", "H_warp_mask"):
        arrays[k + "_sample"] = od[k][..., sy, sx].contiguous().numpy()
    arrays["blend_image_sample"] = od["blend_image"][..., sy, sx].contiguous().numpy()
    for k in ("occlusion_mask", "origin_occlusion_mask", "warp_input2_mask"):
        arrays[k + "_bits"] = np.packbits(od[k].numpy() > 0.5)
        arrays[k + "_shape"] = np.array(od[k].shape)
    arrays["mask1_bits"] = np.packbits(od["mask1"].numpy() > 0.5)
    arrays["mask2_bits"] = np.packbits(od["mask2"].numpy() > 0.5)
    arrays["canvas"] = np.array([od["width_min"], od["height_min"], od["out_height"], od["out_width"]])
    arrays["H"] = od["H"].numpy()
    arrays["I_mat"] = od["I_mat"].numpy()
    save("adapter_test_out", cases.checksum(c["image1"], c["image2"], c["offsets"], *c["flows"]), **arrays)

    # ---------------------------------------------------------------- BASELINE config 1: the reference's
    # demo/demo1 pair through the path on the CPU (out.py:129-146 loading; stub networks per SURVEY 8(d)).
    # Kernel-level vectors (the thetas the reference computed + a pixel subset of every intermediate) so
    # that each CUDA kernel can be checked bit for bit on real image content, plus the adapter's out_dict.
    c = cases.demo1_pair()
    ad = ref_ad.FlowHomoAdpater(cases.StubHomo(c["offsets"]), cases.StubFlow(c["flows"]), cases.adapter_cfg())
    ad.eval()
    calls = dict(transformer=[], warp=[], morph=[], occ=[])
    orig = (ref_ad.torch_homo_transform.transformer, ref_ad.warp, ref_ad.preprocess_occlusion_mask, ref_ad.compute_occlusion)

    def _rec(key, fn):
        def wrapped(*a, **k):
            r = fn(*a, **k)
            calls[key].append((a, k, r.clone()))
            return r
        return wrapped

    ref_ad.torch_homo_transform.transformer = _rec("transformer", orig[0])
    ref_ad.warp = _rec("warp", orig[1])
    ref_ad.preprocess_occlusion_mask = _rec("morph", orig[2])
    ref_ad.compute_occlusion = _rec("occ", orig[3])
    try:
        od = ad.test_out_forward(c["image1"], c["image2"])
    finally:
        (ref_ad.torch_homo_transform.transformer, ref_ad.warp, ref_ad.preprocess_occlusion_mask,
         ref_ad.compute_occlusion) = orig
    assert [len(calls[k]) for k in ("transformer", "warp", "morph", "occ")] == [5, 1, 2, 1]
    dy, dx = cases.DEMO1_SAMPLE
    sub = lambda t: t[..., dy, dx].contiguous().numpy()
    bits = lambda t: np.packbits(t.numpy() > 0.5)
    tr = calls["transformer"]
    arrays = dict(pixels_checksum=np.float64(cases.checksum(c["image1"], c["image2"])),
                  canvas=np.array([od["width_min"], od["height_min"], od["out_height"], od["out_width"]]),
                  theta_H512=tr[0][0][1].numpy(), theta_I=tr[1][0][1].numpy(), theta_H=tr[2][0][1].numpy(),
                  output_H512_sample=sub(tr[0][2]), homo_output_sample=sub(tr[1][2]), homo_output2_sample=sub(tr[2][2]),
                  residual_flow_output_sample=sub(tr[3][2]), warp_out_sample=sub(calls["warp"][0][2]),
                  occ_raw_sample=sub(calls["occ"][0][2]), occ_raw_bits=bits(calls["occ"][0][2]),
                  origin_occlusion_bits=bits(calls["morph"][0][2]), occ_canvas_sample=sub(tr[4][2]),
                  occ_canvas_bits=bits(tr[4][2]), occlusion_bits=bits(calls["morph"][1][2]),
                  warp_input2_mask_bits=bits(od["warp_input2_mask"]), H=od["H"].numpy())
    # at 512^2 resize_flow (:241) is the identity: the flows compute_occlusion saw ARE the case's flows
    assert torch.equal(calls["occ"][0][1]["flow_ij"], c["flows"][0]) and torch.equal(calls["occ"][0][1]["flow_ji"], c["flows"][1])
    for k in ("H_warp", "final_warp", "output1", "output2", "mask1", "mask2", "H_warp_mask", "blend_image"):
        arrays["out_" + k + "_sample"] = sub(od[k])
    arrays["mask1_bits"], arrays["mask2_bits"] = bits(od["mask1"]), bits(od["mask2"])
    # W3 on real pixels: the UDIS TPS warp of (image1 | ones) with the 13x13 mesh
    U = torch.cat((c["image1"], torch.ones_like(c["image1"])), 1)
    tps_out, rec = capture_gather_indices(ref_tps.transformer, U, c["tps_source"], c["tps_target"], (512, 512))
    arrays["tps_out_sample"] = sub(tps_out)
    arrays["tps_idx_sample"] = decode_indices(rec, 1, 512, 512, 512, 512)[..., dy, dx]
    save("demo1_pair", cases.checksum(c["image1"], c["image2"], c["offsets"], *c["flows"], c["tps_source"], c["tps_target"]),
         **arrays)


if __name__ == "__main__":
    main()
