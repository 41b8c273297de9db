"""CPU: the oracle (oracle/oracle.c + stitch_oracle.py) against golden vectors
produced by the reference's own functions (tests/golden/make_golden.py).
This is what pins the oracle; the GPU tests then compare the kernels with it."""
import numpy as np
import pytest
import torch

import cases
import stitch_oracle as so
from conftest import golden
from helpers import assert_bits_equal, check_inputs, max_abs, unpack_bits


# ---------------------------------------------------------------- C1 / C2
@pytest.mark.parametrize("name", ["corr_small", "corr_c64"])
def test_corr_small(name):
    c = getattr(cases, name)()
    g = golden(name)
    check_inputs(g, *c.values())
    for blas in (True, False):
        vol = so.corr(c["fmap1"].numpy(), c["fmap2"].numpy(), use_blas=blas)
        # fp32 contraction, K = 256: summation order differs between BLAS builds -> 1e-4 abs on |v| ~ 16
        assert max_abs(vol, g["vol"]) < 2e-4


def test_corr_512_sample_and_pyramid():
    c = cases.corr_512()
    g = golden("corr_512")
    check_inputs(g, *c.values())
    vol = so.corr(c["fmap1"].numpy(), c["fmap2"].numpy())
    v2 = vol.reshape(4096, 4096)
    assert max_abs(v2[cases.CORR_512_ROWS, cases.CORR_512_COLS], g["vol_sample"]) < 3e-4
    pyr = so.corr_pyramid(c["fmap1"].numpy(), c["fmap2"].numpy(), 4)
    # C2 has no reference implementation: pinned against torch's avg_pool2d chain
    for l, key in ((1, "lvl1_sample"), (2, "lvl2_sample"), (3, "lvl3_sample")):
        assert max_abs(pyr[l][::97], g[key]) < 3e-4


def test_avg_pool_bit_exact_vs_torch():
    x = torch.randn(5, 1, 16, 24, generator=torch.Generator().manual_seed(3))
    ref = torch.nn.functional.avg_pool2d(x, 2, stride=2).numpy()
    assert_bits_equal(so.avg_pool2x2(x.numpy()), ref, "avg_pool2x2")


# ---------------------------------------------------------------- C3
@pytest.mark.parametrize("name", ["lookup_small", "lookup_64"])
def test_lookup(name):
    c = getattr(cases, name)()
    g = golden(name)
    check_inputs(g, *c.values())
    out = so.encode_flow_token(c["cost_maps"].numpy(), c["coords"].numpy())
    # same fp32 op sequence as ATen's CPU grid_sample: bit equality
    assert_bits_equal(out, g["out"], "lookup")
    assert tuple(g["out_strides"]) == tuple(s // 4 for s in out.strides)
    if name == "lookup_small":
        for rr in (0, 1, 2, 7):
            assert_bits_equal(so.encode_flow_token(c["cost_maps"].numpy(), c["coords"].numpy(), r=rr), g[f"out_r{rr}"], f"r={rr}")
        pyr = [c["cost_maps"].numpy()]
        for _ in range(2):
            pyr.append(so.avg_pool2x2(pyr[-1]))
        assert_bits_equal(so.encode_flow_token_pyramid(pyr, c["coords"].numpy()), g["out_pyramid"], "pyramid lookup")


def test_bilinear_sampler():
    g = golden("bilinear_sampler")
    assert_bits_equal(so.bilinear_sampler(g["img"], g["pts"]), g["out"], "bilinear_sampler")


# ---------------------------------------------------------------- W1
@pytest.mark.parametrize("name", ["warp_small", "warp_flow2"])
def test_warp(name):
    c = getattr(cases, name)()
    g = golden(name)
    check_inputs(g, *c.values())
    out = so.warp(c["x"].numpy(), c["flo"].numpy())
    # contract: 1e-3 max-abs on 0..255 images; the restated arithmetic is bit-exact
    assert_bits_equal(out, g["out"], "warp")
    # mode='nearest' (warp_utils.py:76-77 -> grid_sample nearest, half-to-even rounding, zeros outside)
    assert_bits_equal(so.warp(c["x"].numpy(), c["flo"].numpy(), mode="nearest"), g["out_nearest"], "warp, nearest")


def test_warp_512():
    c = cases.warp_512()
    g = golden("warp_512")
    check_inputs(g, *c.values())
    out = so.warp(c["x"].numpy(), c["flo"].numpy())
    sy, sx = cases.WARP_512_SAMPLE
    assert_bits_equal(out[..., sy, sx], g["out_sample"], "warp 512")


# ---------------------------------------------------------------- W2
@pytest.mark.parametrize("name", ["homo_small", "homo_theta1", "homo_degenerate"])
def test_homo(name):
    c = getattr(cases, name)()
    g = golden(name)
    check_inputs(g, c["U"], c["theta"])
    out, idx = so.homo_transformer(c["U"].numpy(), c["theta"].numpy(), c["out_size"], return_indices=True)
    assert_bits_equal(idx, g["idx"], "integer grid indices")          # contract: bit-exact
    if name == "homo_degenerate":
        fin = np.isfinite(g["out"]) & (np.abs(g["out"]) < 1e6)
        assert max_abs(np.where(fin, out, 0), np.where(fin, g["out"], 0)) <= 1e-3
    else:
        assert_bits_equal(out, g["out"], "warped values")               # same op order -> identical bits


def test_homo_512():
    c = cases.homo_512()
    g = golden("homo_512")
    check_inputs(g, c["U"], c["theta"])
    out, idx = so.homo_transformer(c["U"].numpy(), c["theta"].numpy(), c["out_size"], return_indices=True)
    sy, sx = cases.HOMO_512_SAMPLE
    assert_bits_equal(idx[..., sy, sx], g["idx_sample"], "integer grid indices")
    assert_bits_equal(out[..., sy, sx], g["out_sample"], "warped values")
    assert_bits_equal(np.packbits(out[0, 3] > 0.5), g["mask_bits"], "thresholded mask")


# ---------------------------------------------------------------- W3
def test_tps():
    c = cases.tps_small()
    g = golden("tps_small")
    check_inputs(g, c["U"], c["source"], c["target"])
    out, idx = so.tps_transformer(c["U"].numpy(), c["source"].numpy(), c["target"].numpy(), c["out_size"],
                                  return_indices=True)
    # BLAS summation order of the reference's T @ basis is unspecified: coordinates agree to
    # ~1e-5, so a few samples sitting on an integer boundary may floor differently
    mism = (idx != g["idx"]).any(axis=1)
    assert mism.mean() < 0.01
    ok = ~mism[:, None].repeat(6, 1)
    assert max_abs(np.where(ok, out, 0), np.where(ok, g["out"], 0)) <= 1e-3


# ---------------------------------------------------------------- N1
def test_gma_attention_and_aggregate():
    c = cases.gma_small()
    g = golden("gma_small")
    check_inputs(g, *c.values())
    attn = so.gma_attention(c["fmap"].numpy(), c["w_qk"].numpy())
    assert max_abs(attn, g["attn"]) <= 5e-6
    out = so.gma_aggregate(g["attn"], c["motion"].numpy(), c["w_v"].numpy(), float(c["gamma"]))
    assert max_abs(out, g["out"]) <= 5e-6
    # the tensor-core contract: bf16-rounded q, k move the probabilities by < 2 % of the row maximum
    attn_bf = so.gma_attention(c["fmap"].numpy(), c["w_qk"].numpy(), bf16_inputs=True)
    assert (np.abs(attn_bf - g["attn"]).max(-1) / g["attn"].max(-1)).max() <= 2e-2


# ---------------------------------------------------------------- N3
def test_ccl():
    c = cases.ccl_small()
    g = golden("ccl_small")
    check_inputs(g, *c.values())
    flow = so.ccl(c["feature_1"].numpy(), c["feature_2"].numpy())
    # fp64 restatement (nine shifted diagonals of the plain correlation) vs the reference's fp32 conv2d
    assert max_abs(flow, g["flow"]) <= 5e-5


# ---------------------------------------------------------------- N2
def test_upsample_flow():
    c = cases.upsample_small()
    g = golden("upsample_small")
    check_inputs(g, *c.values())
    out = so.upsample_flow(c["flow"].numpy(), c["mask"].numpy())
    # same op sequence; ATen's vectorised exp (Sleef) and its 9-term sums differ from libm / sequential
    # order by a few ulp: 1.1e-5 on |v| up to 82
    assert max_abs(out, g["out"]) <= 3e-5


# ---------------------------------------------------------------- W3k
def test_tps_kornia():
    """warp_image_tps of the reference's kornia_tps.py (warp_points_tps / create_meshgrid restated
    from kornia, which is absent and unpinned: 'pinned modulo that restatement')."""
    c = cases.tps_kornia_small()
    g = golden("tps_kornia")
    check_inputs(g, *c.values())
    img = c["image"].numpy()
    # the sampler is pinned exactly: reference grid -> reference output, both align_corners modes
    for ac in (0, 1):
        assert_bits_equal(so.grid_sample(img, g["grid"], bool(ac)), g[f"out_ac{ac}"], f"grid_sample align_corners={ac}")
    # the K-term fp32 sum of the reference is order-unspecified (torch cascade sum): grid to ~1e-6
    grid = so.tps_kornia_grid(c["points_src"].numpy(), g["kernel_weights"], g["affine_weights"], img.shape[2], img.shape[3])
    assert max_abs(grid, g["grid"]) <= 5e-6
    for ac in (0, 1):
        out = so.warp_image_tps(img, c["points_src"].numpy(), g["kernel_weights"], g["affine_weights"], bool(ac))
        assert max_abs(out, g[f"out_ac{ac}"]) <= 1e-3


# ---------------------------------------------------------------- W4
@pytest.mark.parametrize("name", ["range_small", "range_smooth"])
def test_range_map(name):
    c = getattr(cases, name)()
    g = golden(name)
    check_inputs(g, *c.values())
    rm = so.compute_range_map(c["flow_ji"].numpy())
    assert_bits_equal(rm, g["range_map"], "range map (same sequential fp32 order as CPU scatter_add_)")
    assert_bits_equal(so.compute_occlusion_wang(c["flow_ji"].numpy(), True), g["occ"], "occlusion")
    assert_bits_equal(so.compute_occlusion_wang(c["flow_ji"].numpy(), False), g["occ_nz"], "occlusion (ones)")
    assert_bits_equal(so.compute_occlusion_wang(c["flow_ji"].numpy(), True, threshold=True),
                      (g["occ"] >= 0.5).astype(np.float32), "thresholded occlusion")


# ---------------------------------------------------------------- W5
@pytest.mark.parametrize("name", ["morph_small", "morph_big"])
def test_morph(name):
    c = getattr(cases, name)()
    g = golden(name)
    check_inputs(g, c["mask"])
    shape = tuple(g["shape"])
    out = so.preprocess_occlusion_mask(c["mask"].numpy())
    assert_bits_equal(out > 0.5, unpack_bits(g["out_bits"], shape), "19x19 open")
    out7 = so.preprocess_occlusion_mask(c["mask"].numpy(), (7, 11))
    assert_bits_equal(out7 > 0.5, unpack_bits(g["out7_bits"], shape), "7x11 open")
    assert set(np.unique(out)) <= {0.0, 1.0}


# ---------------------------------------------------------------- W6 / W7 / W8
def test_composite_w6_reference_block():
    """W6: the oracle's fused compositing against the inputs and outputs of the reference's OWN block
    (flowHomoAdpater.py:317,339-360), captured inside a run of its test_out_forward (make_golden.py spies on
    transformer / warp / preprocess_occlusion_mask).  Per-pixel arithmetic, so the stored pixel subset
    (cases.W6_SAMPLE) is a complete vector.  Bit-exact, including the uint8 blend and the 0/0 pixels."""
    g = golden("composite_w6")
    fw_in = g["warp_out"] * g["flow_mask"]                                   # :317 (one fp32 multiply)
    assert_bits_equal(fw_in.view(np.uint32), g["final_warp_in"].view(np.uint32), "final_warp * flow_mask")
    r = so.composite_test_out(g["homo_output"], g["homo_output2"], fw_in, g["occlusion_mask"])
    assert_bits_equal(r["final_warp_output"][:, 0:3].view(np.uint32), g["out_final_warp"].view(np.uint32), "final_warp")
    for k in ("output1", "output2", "mask1", "mask2"):
        a, b = np.ascontiguousarray(r[k]), g["out_" + k]
        nan_a, nan_b = np.isnan(a), np.isnan(b)
        assert_bits_equal(nan_a, nan_b, k + " NaN pattern")
        assert_bits_equal(a.view(np.uint32)[~nan_a], b.view(np.uint32)[~nan_b], k)
    assert r["blend_image"].dtype == np.uint8
    assert_bits_equal(r["blend_image"], g["out_blend_image"], "blend_image (uint8)")
    # the case exercises every region: overlap, img1 only, img2 only, nothing, occluded
    m1, m2 = g["out_mask1"][0, 0] > 0.5, g["out_mask2"][0, 0] > 0.5
    for region in (m1 & m2, m1 & ~m2, ~m1 & m2, ~m1 & ~m2, g["occlusion_mask"][0, 0] < 0.5):
        assert region.sum() > 20


def test_build_model():
    c = cases.build_model_small()
    g = golden("build_model")
    check_inputs(g, *c.values())
    r = so.build_model_arith(*[c[k].numpy() for k in ("warp1", "warp2", "mask1", "mask2", "net_out")])
    for k in ("learned_mask1", "learned_mask2", "stitched_image"):
        assert_bits_equal(r[k], g[k], k)


def test_tps_mix():
    """W8 against the reference's own tps_H_warp (core/inference/tps_pipline.py:138-170 executed as written,
    stages upstream of :138 stubbed — see make_golden.py)."""
    c = cases.tps_mix_small()
    g = golden("tps_mix")
    check_inputs(g, *c.values())
    tm3 = c["tps_mask3"].numpy()
    tm = (tm3.mean(axis=1, keepdims=True) >= 0.5).astype(np.float32)
    tm = 1.0 - so.morph_open(1.0 - tm, (11, 11), border_is_zero=False)
    assert_bits_equal(tm, g["tps_mask"], "11x11 cv2 open of the inverse mask")
    out2, mask2, blend = so.tps_mix_blend(c["final_warp"].numpy(), c["tps_warp_raw"].numpy() * tm, tm,
                                          c["output1"].numpy(), c["mask1"].numpy())
    assert_bits_equal(out2, g["output2"], "output2")
    assert_bits_equal(mask2, g["mask2"], "mask2")
    assert_bits_equal(blend, g["blend"], "blend")


# ---------------------------------------------------------------- N4
def test_patch_embed_proj():
    """The oracle's conv stack against the reference's own PatchEmbed (output of its last conv, captured while the
    module's forward ran): fp32 summation-order noise only.  The bf16-operand emulation of the tensor-core kernel
    stays within the 1e-2-of-scale contract of the fp32 reference."""
    c = cases.patch_embed_small()
    g = golden("patch_embed_small")
    check_inputs(g, *c.values())
    a = {k: v.numpy() for k, v in c.items()}
    scale = float(np.abs(g["proj"]).max())
    out = so.patch_embed_proj(**a)
    assert out.shape == g["proj"].shape == (6, 64, 8, 8)
    assert max_abs(out, g["proj"]) <= 2e-5 * scale
    assert max_abs(so.patch_embed_proj(**a, bf16_operands=True), g["proj"]) <= 1e-2 * scale
    # the all-zero map: every layer sees only its bias through the zero padding pattern
    assert max_abs(out[5, :, 3, 3], so.patch_embed_proj(np.zeros((1, 1, 64, 64), np.float32), *[a[k] for k in ("w1", "b1", "w2", "b2", "w3", "b3")])[0, :, 3, 3]) == 0.0


# ---------------------------------------------------------------- G1
def test_dlt():
    g = golden("geometry")
    src = np.tile(np.array([[0.0, 0.0], [64, 0.0], [0.0, 48], [64, 48]], np.float32)[None], (3, 1, 1))
    H = so.tensor_DLT(src, g["dst"])
    assert max_abs(H, g["H"]) < 1e-4
