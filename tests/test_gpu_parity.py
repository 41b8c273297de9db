"""GPU (B200): the CUDA kernels, called through the reference-shaped Python
surface -> ctypes -> C ABI, against (a) the CPU oracle on the same seeded inputs
and (b) the committed golden vectors of the reference itself.

Tolerances (BASELINE.json north_star): integer grid indices and (thresholded)
masks bit-exact; cost volume within 1e-2 relative to the volume's scale (bf16
contraction vs fp32 reference); warped images within 1e-3 max-abs — the gathers
restate the oracle's fp32 op order and are checked for bit equality where that
order is fully specified.
"""
import numpy as np
import pytest
import torch

import cases
import stitch_oracle as so
from conftest import golden
from helpers import assert_bits_equal, check_inputs, max_abs, unpack_bits

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sb():
    if not torch.cuda.is_available():
        pytest.fail("gpu tests need a CUDA device")
    import stitch_b200
    assert stitch_b200._lib.load().sb_device_check() == 0, stitch_b200._lib.last_error()
    return stitch_b200


def cu(t):
    return t.cuda() if isinstance(t, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(t)).cuda()


def host(t):
    torch.cuda.synchronize()
    return t.detach().cpu().numpy()


# ===================================================================== C1 / C2
def _corr_check(vol_gpu, f1, f2):
    """bf16-operand contraction vs (1) the same contraction on bf16-rounded inputs in
    fp64 (kernel correctness, tight) and (2) the reference's fp32 result (the contract)."""
    ref32 = so.corr(f1, f2)
    ref_bf = so.corr_bf16_inputs(f1, f2)
    scale = float(np.abs(ref32).max())
    rms = float(np.sqrt((ref32.astype(np.float64) ** 2).mean()))
    err_kernel = max_abs(vol_gpu, ref_bf)
    err_contract = max_abs(vol_gpu, ref32)
    fro = float(np.linalg.norm((vol_gpu - ref32).ravel()) / max(np.linalg.norm(ref32.ravel()), 1e-30))
    assert err_kernel <= 2e-5 * max(scale, 1.0), f"kernel error {err_kernel} (scale {scale})"
    assert err_contract <= 1e-2 * scale, f"contract: max-abs {err_contract} vs 1e-2 * {scale}"
    assert fro <= 1e-2, f"contract: relative Frobenius error {fro}"
    return err_kernel, err_contract, rms


def test_corr_bf16_volume_option(sb):
    """Opt-in bf16 volume: equals the fp32 volume of the same kernel rounded to bf16 (round-to-nearest-even)."""
    gen = torch.Generator().manual_seed(10)
    f1 = torch.randn(2, 256, 16, 24, generator=gen)
    f2 = torch.randn(2, 256, 16, 24, generator=gen)
    t1, t2 = sb.corr.tokens_bf16(cu(f1)), sb.corr.tokens_bf16(cu(f2))
    v32 = sb.corr.corr_from_tokens(t1, t2, 256, (16, 24), (16, 24))
    v16 = sb.corr.corr_from_tokens(t1, t2, 256, (16, 24), (16, 24), out_dtype=torch.bfloat16)
    assert v16.dtype == torch.bfloat16 and v16.shape == v32.shape
    assert torch.equal(v16, v32.bfloat16())
    ref = so.corr(f1.numpy(), f2.numpy())
    assert max_abs(host(v16.float()), ref) <= 1e-2 * float(np.abs(ref).max())


def test_corr_cta_pair_mode_is_bit_identical(sb):
    """tcgen05 cta_group::2 variant (opt-in through sb_tune): same bits as the one-CTA-per-tile kernel,
    volume and fused pyramid, ragged and full shapes."""
    lib = sb._lib.load()
    gen = torch.Generator().manual_seed(14)
    try:
        for b, hw, lv in ((1, (9, 20), 0), (2, (64, 64), 3), (1, (24, 40), 0)):
            f1 = torch.randn(b, 256, *hw, generator=gen)
            f2 = torch.randn(b, 256, *hw, generator=gen)
            t1, t2 = sb.corr.tokens_bf16(cu(f1)), sb.corr.tokens_bf16(cu(f2))
            res = []
            for mode in (1, 2):
                lib.sb_tune(6, mode)
                res.append(sb.corr.corr_from_tokens(t1, t2, 256, hw, hw, pyramid_levels=lv))
            if lv:
                assert torch.equal(res[0][0], res[1][0])
                for x, y in zip(res[0][1], res[1][1]):
                    assert torch.equal(x, y)
            else:
                assert torch.equal(res[0], res[1])
    finally:
        lib.sb_tune(6, 0)


def test_corr_dynamic_units_bit_identical(sb):
    """Work units handed out by the hardware scheduler (clusterlaunchcontrol.try_cancel, the default) vs the static
    round-robin split (sb_tune 14 = 2; 0 means "the default"): same bits — volume, fused pyramid (64-wide and the 128-wide pair-of-tiles
    epilogue), the two-pass attention logits — on shapes with more units than SMs."""
    lib = sb._lib.load()
    gen = torch.Generator().manual_seed(15)
    try:
        for b, hw, lv in ((3, (64, 64), 3), (2, (64, 64), 0), (1, (128, 128), 3), (5, (40, 56), 0)):
            f1 = torch.randn(b, 256, *hw, generator=gen)
            f2 = torch.randn(b, 256, *hw, generator=gen)
            t1, t2 = sb.corr.tokens_bf16(cu(f1)), sb.corr.tokens_bf16(cu(f2))
            res = []
            for mode in (2, 1):
                assert lib.sb_tune(14, mode) == 0
                res.append(sb.corr.corr_from_tokens(t1, t2, 256, hw, hw, pyramid_levels=lv))
            if lv:
                assert torch.equal(res[0][0], res[1][0])
                for x, y in zip(res[0][1], res[1][1]):
                    assert torch.equal(x, y)
            else:
                assert torch.equal(res[0], res[1])
        fmap = cu(torch.randn(2, 128, 64, 64, generator=gen))
        w_qk = cu(torch.randn(256, 128, 1, 1, generator=gen) * 0.05)
        att = []
        for mode in (2, 1):
            assert lib.sb_tune(14, mode) == 0
            att.append(sb.gma.attention(fmap, w_qk, heads=1))
        assert torch.equal(att[0], att[1])
    finally:
        lib.sb_tune(14, 0)


def test_corr_a_operand_from_tensor_memory_is_bit_identical(sb):
    """A block copied to TMEM once per unit (tcgen05.cp) and read by the TS form of tcgen05.mma (the default;
    sb_tune 10 = 2 selects shared-memory operands): same bits as the shared-memory-operand kernel; ragged shapes,
    C = 96 (padded K) and 256."""
    lib = sb._lib.load()
    gen = torch.Generator().manual_seed(15)
    try:
        for b, c, hw, lv in ((1, 256, (9, 20), 0), (2, 256, (64, 64), 3), (1, 96, (24, 40), 0), (3, 128, (20, 32), 0)):
            f1 = torch.randn(b, c, *hw, generator=gen)
            f2 = torch.randn(b, c, *hw, generator=gen)
            t1, t2 = sb.corr.tokens_bf16(cu(f1)), sb.corr.tokens_bf16(cu(f2))
            res = []
            for mode in (2, 1):
                lib.sb_tune(10, mode)
                res.append(sb.corr.corr_from_tokens(t1, t2, c, hw, hw, pyramid_levels=lv))
            if lv:
                assert torch.equal(res[0][0], res[1][0])
                for x, y in zip(res[0][1], res[1][1]):
                    assert torch.equal(x, y)
            else:
                assert torch.equal(res[0], res[1])
    finally:
        lib.sb_tune(10, 0)


def test_corr_odd_token_count(sb):
    """13 x 15 = 195 target tokens (not a multiple of 4): pitched volume, strided view of the reference's shape."""
    gen = torch.Generator().manual_seed(9)
    f1 = torch.randn(2, 96, 9, 11, generator=gen)
    f2 = torch.randn(2, 96, 13, 15, generator=gen)
    vol = sb.corr.corr(cu(f1), cu(f2))
    assert tuple(vol.shape) == (2, 1, 9, 11, 13, 15)
    _corr_check(host(vol.contiguous()), f1.numpy(), f2.numpy())
    # and the lookup on the (copied-contiguous) maps of that volume, generic-width path
    maps = vol.contiguous().view(2 * 99, 1, 13, 15)
    coords = sb.lookup.coords_grid(2, 9, 11, device="cuda") + torch.randn(2, 2, 9, 11, device="cuda") * 2
    out = sb.encode_flow_token(maps, coords)
    assert_bits_equal(host(out.contiguous()), np.ascontiguousarray(so.encode_flow_token(host(maps), host(coords))), "lookup W2=15")


@pytest.mark.parametrize("name", ["corr_c64", "corr_small"])
def test_corr_small_cases(sb, name):
    c = getattr(cases, name)()
    g = golden(name)
    check_inputs(g, *c.values())
    vol = host(sb.corr.corr(cu(c["fmap1"]), cu(c["fmap2"])))
    assert vol.shape == g["vol"].shape
    _corr_check(vol, c["fmap1"].numpy(), c["fmap2"].numpy())
    assert max_abs(vol, g["vol"]) <= 1e-2 * float(np.abs(g["vol"]).max())


@pytest.mark.parametrize("shape", [(1, 64, 8, 16, 8, 16), (1, 128, 12, 11, 9, 12), (3, 192, 16, 16, 20, 16),
                                   (2, 40, 5, 7, 6, 6), (1, 256, 16, 16, 32, 64)])
def test_corr_ragged_shapes(sb, shape):
    """tails in M, N and K (TMA zero fill / clipped stores), several batches and panels."""
    b, ch, h1, w1, h2, w2 = shape
    g = torch.Generator().manual_seed(sum(shape))
    f1 = torch.randn(b, ch, h1, w1, generator=g)
    f2 = torch.randn(b, ch, h2, w2, generator=g)
    vol = host(sb.corr.corr(cu(f1), cu(f2)))
    assert vol.shape == (b, 1, h1, w1, h2, w2)
    _corr_check(vol, f1.numpy(), f2.numpy())


def test_corr_heads(sb):
    g = torch.Generator().manual_seed(5)
    f1 = torch.randn(2, 128, 8, 8, generator=g)
    f2 = torch.randn(2, 128, 8, 8, generator=g)
    vol = host(sb.corr.corr(cu(f1), cu(f2), heads=2))
    ref = so.corr(f1.numpy(), f2.numpy(), heads=2)
    assert vol.shape == ref.shape == (2, 2, 8, 8, 8, 8)
    assert max_abs(vol, ref) <= 1e-2 * float(np.abs(ref).max())


def test_corr_512_golden_and_pyramid(sb):
    c = cases.corr_512()
    g = golden("corr_512")
    check_inputs(g, *c.values())
    vol, lv = sb.corr.corr(cu(c["fmap1"]), cu(c["fmap2"]), pyramid_levels=3)
    v = host(vol).reshape(4096, 4096)
    scale = float(g["vol_absmax"])
    assert max_abs(v[cases.CORR_512_ROWS, cases.CORR_512_COLS], g["vol_sample"]) <= 1e-2 * scale
    _corr_check(host(vol), c["fmap1"].numpy(), c["fmap2"].numpy())
    # fused pyramid == chained avg_pool2d of the volume the kernel itself wrote (bit-exact: same
    # ((a+b)+c)+d order, * 0.25) and within the contract of the reference-derived golden levels
    cm = v.reshape(4096, 1, 64, 64)
    l1 = so.avg_pool2x2(cm); l2 = so.avg_pool2x2(l1); l3 = so.avg_pool2x2(l2)
    for got, want, key in ((lv[0], l1, "lvl1_sample"), (lv[1], l2, "lvl2_sample"), (lv[2], l3, "lvl3_sample")):
        got = host(got)
        assert got.shape == want.shape
        assert_bits_equal(got, want, key + " vs pooled own volume")
        assert max_abs(got[::97], g[key]) <= 1e-2 * scale
    # public pyramid API
    pyr = sb.corr_pyramid(cu(c["fmap1"]), cu(c["fmap2"]), 4)
    assert [tuple(p.shape) for p in pyr] == [(4096, 1, 64, 64), (4096, 1, 32, 32), (4096, 1, 16, 16), (4096, 1, 8, 8)]
    assert torch.equal(pyr[0].view(-1), vol.view(-1))      # level 0 of the public pyramid IS the volume
    for p_, l_ in zip(pyr[1:], lv):
        assert torch.equal(p_.view(-1), l_.view(-1))


def test_corr_pyramid_generic_width(sb):
    """W2 != 64: the pyramid falls back to the standalone pooling kernel."""
    g = torch.Generator().manual_seed(9)
    f1 = torch.randn(1, 64, 6, 6, generator=g)
    f2 = torch.randn(1, 64, 16, 24, generator=g)
    vol, lv = sb.corr.corr(cu(f1), cu(f2), pyramid_levels=3)
    cm = host(vol).reshape(36, 1, 16, 24)
    l1 = so.avg_pool2x2(cm); l2 = so.avg_pool2x2(l1); l3 = so.avg_pool2x2(l2)
    for got, want in zip(lv, (l1, l2, l3)):
        assert_bits_equal(host(got), want, "standalone pooling")


def test_corr_full_batch_properties(sb):
    """BASELINE config 2 size (B=16, 512^2 -> N=4096): size-independent properties."""
    g = torch.Generator(device="cuda").manual_seed(1)
    f1 = torch.randn(16, 256, 64, 64, device="cuda", generator=g)
    f2 = torch.randn(16, 256, 64, 64, device="cuda", generator=g)
    vol = sb.corr.corr(f1, f2).view(16, 4096, 4096)
    # torch fp32 reference of the same op on bf16-rounded operands (floating-point kernel)
    ref = torch.bmm(f1.bfloat16().float().view(16, 256, 4096).transpose(1, 2), f2.bfloat16().float().view(16, 256, 4096))
    scale = ref.abs().max().item()
    assert (vol - ref).abs().max().item() <= 5e-5 * scale
    ref32 = torch.bmm(f1.view(16, 256, 4096).transpose(1, 2), f2.view(16, 256, 4096))
    assert (vol - ref32).abs().max().item() <= 1e-2 * ref32.abs().max().item()
    del ref, ref32
    # transpose symmetry: corr(f2, f1)[b] == corr(f1, f2)[b]^T
    vol_t = sb.corr.corr(f2, f1).view(16, 4096, 4096)
    assert (vol_t - vol.transpose(1, 2)).abs().max().item() <= 1e-5 * scale
    del vol_t
    # exact linearity under power-of-two scaling
    vol2 = sb.corr.corr(f1 * 2.0, f2 * 0.5).view(16, 4096, 4096)
    assert torch.equal(vol2, vol)


def test_corr_bidirectional_is_both_volumes_bit_for_bit(sb):
    """Forward and backward volume (+ fused pyramid) of a batch of pairs from ONE launch == the two separate launches."""
    g = torch.Generator(device="cuda").manual_seed(5)
    for (b, c, h, w, lv) in ((3, 256, 64, 64, 3), (2, 96, 16, 24, 0), (1, 256, 64, 64, 0), (5, 64, 8, 16, 3)):
        f1 = torch.randn(b, c, h, w, device="cuda", generator=g)
        f2 = torch.randn(b, c, h, w, device="cuda", generator=g)
        res = sb.corr.corr_bidirectional(f1, f2, pyramid_levels=lv)
        fwd = sb.corr.corr(f1, f2, pyramid_levels=lv)
        bwd = sb.corr.corr(f2, f1, pyramid_levels=lv)
        vol = res[0] if lv else res
        vf, vb = (fwd[0], bwd[0]) if lv else (fwd, bwd)
        assert vol.shape == (2 * b, 1, h, w, h, w)
        assert torch.equal(vol[:b], vf) and torch.equal(vol[b:], vb)
        for l in range(lv):
            both = res[1][l].view(2, -1)
            assert torch.equal(both[0], fwd[1][l].view(-1)) and torch.equal(both[1], bwd[1][l].view(-1))
    with pytest.raises(ValueError):
        sb.corr.corr_bidirectional(torch.zeros(1, 64, 8, 8, device="cuda"), torch.zeros(1, 64, 8, 16, device="cuda"))


def test_step_bidirectional_equals_two_directions(sb):
    """HotPath with the two directions batched into one launch each (21 launches) == the per-direction step (34)."""
    from stitch_b200.pipeline import HotPath, make_pair_batch
    pb = make_pair_batch(3, 3, size=256, iters=3).map(lambda t: t.cuda())
    a = HotPath(size=256, iters=3, pyramid=True, bidirectional=True).step(pb)
    sb._lib.reset_launch_count()
    b_ = HotPath(size=256, iters=3, pyramid=True, bidirectional=False).step(pb)
    torch.cuda.synchronize()
    for k in ("cost_volume", "cost_volume_back", "final_warp_output", "overlap", "origin_occlusion_mask"):
        assert torch.equal(a[k], b_[k]), k
    assert len(a["cost_tokens"]) == len(b_["cost_tokens"]) == 6
    for x, y in zip(a["cost_tokens"], b_["cost_tokens"]):
        assert torch.equal(x, y)
    for x, y in zip(a["cost_pyramid"] + a["cost_pyramid_back"], b_["cost_pyramid"] + b_["cost_pyramid_back"]):
        assert torch.equal(x, y)


def test_step_two_lookup_chains_equal_one(sb):
    """The forward / backward lookup chains on two streams (default) == one chain after the other."""
    from stitch_b200.pipeline import HotPath, make_pair_batch
    pb = make_pair_batch(5, 3, size=256, iters=4).map(lambda t: t.cuda())
    outs = []
    for n in (1, 2):
        hp = HotPath(size=256, iters=4, pyramid=True)
        hp.lookup_streams = n
        outs.append(hp.step(pb))
        torch.cuda.synchronize()
    assert len(outs[0]["cost_tokens"]) == len(outs[1]["cost_tokens"]) == 8
    for x, y in zip(outs[0]["cost_tokens"], outs[1]["cost_tokens"]):
        assert torch.equal(x, y)
    for k in ("cost_volume", "cost_volume_back", "final_warp_output"):
        assert torch.equal(outs[0][k], outs[1][k]), k


def test_corr_errors(sb):
    with pytest.raises(RuntimeError, match="C=320"):
        sb.corr.corr(torch.zeros(1, 320, 8, 8, device="cuda"), torch.zeros(1, 320, 8, 8, device="cuda"))
    z = sb.corr.corr(torch.zeros(0, 64, 8, 8, device="cuda"), torch.zeros(0, 64, 8, 8, device="cuda"))
    assert z.shape == (0, 1, 8, 8, 8, 8)


# ===================================================================== C3 / C3p
@pytest.mark.parametrize("name", ["lookup_small", "lookup_64"])
def test_lookup_cases(sb, name):
    c = getattr(cases, name)()
    g = golden(name)
    check_inputs(g, *c.values())
    out = sb.encode_flow_token(cu(c["cost_maps"]), cu(c["coords"]))
    assert tuple(out.stride()) == tuple(int(s) for s in g["out_strides"])   # [B,H1,W1,81] memory order
    assert_bits_equal(host(out.contiguous()), g["out"], "lookup vs reference golden")
    assert_bits_equal(host(out.contiguous()), np.ascontiguousarray(so.encode_flow_token(c["cost_maps"].numpy(), c["coords"].numpy())), "lookup vs oracle")
    if name == "lookup_small":
        out2 = sb.encode_flow_token(cu(c["cost_maps"]), cu(c["coords"]), r=2)
        assert_bits_equal(host(out2.contiguous()), g["out_r2"], "r=2")
        pyr = [cu(c["cost_maps"])]
        for _ in range(2):
            pyr.append(torch.nn.functional.avg_pool2d(pyr[-1], 2, stride=2))
        outp = sb.encode_flow_token_pyramid(pyr, cu(c["coords"]))
        assert max_abs(host(outp.contiguous()), g["out_pyramid"]) <= 1e-5    # pooled maps: torch-GPU pooling order


def test_lookup_non_finite_and_extreme_coords(sb):
    """NaN / +-inf / huge / denormal centres: same values as the oracle (NaN where it is NaN)."""
    gen = torch.Generator().manual_seed(12)
    b, h1, w1 = 1, 4, 8
    maps = torch.randn(b * h1 * w1, 1, 64, 64, generator=gen)
    coords = torch.rand(b, 2, h1, w1, generator=gen) * 60
    special = [float("nan"), float("inf"), -float("inf"), 1e30, -1e30, 1e-40, -1e-40, 3e9, -3e9, 63.0, 0.0, -0.0]
    for i, v in enumerate(special):
        coords[0, i % 2, (i // 8) % h1, i % w1] = v
    out = host(sb.encode_flow_token(cu(maps), cu(coords)).contiguous())
    ref = np.ascontiguousarray(so.encode_flow_token(maps.numpy(), coords.numpy()))
    assert np.array_equal(np.isnan(out), np.isnan(ref))
    assert np.array_equal(np.nan_to_num(out, nan=0.0), np.nan_to_num(ref, nan=0.0))


def test_flow_warp_non_finite_flow(sb):
    gen = torch.Generator().manual_seed(13)
    x = torch.rand(1, 6, 24, 40, generator=gen) * 255
    flo = torch.randn(1, 2, 24, 40, generator=gen) * 3
    for i, v in enumerate([float("nan"), float("inf"), -float("inf"), 1e30, -1e30, 1e-40, 3e9]):
        flo[0, i % 2, 3 + i, 5 + 2 * i] = v
    out = host(sb.warp(cu(x), cu(flo)))
    ref = so.warp(x.numpy(), flo.numpy())
    assert np.array_equal(np.isnan(out), np.isnan(ref))
    assert np.array_equal(np.nan_to_num(out, nan=0.0), np.nan_to_num(ref, nan=0.0))


def test_lookup_full_size_vs_oracle(sb):
    """B=16 x 4096 queries on 64x64 maps (config 2); the oracle checks two of the batches."""
    g = torch.Generator(device="cuda").manual_seed(2)
    b = 16
    maps = torch.randn(b * 4096, 1, 64, 64, device="cuda", generator=g)
    coords = sb.lookup.coords_grid(b, 64, 64, device="cuda") + torch.randn(b, 2, 64, 64, device="cuda", generator=g) * 3.0
    coords[:, :, 0, :] = sb.lookup.coords_grid(b, 64, 64, device="cuda")[:, :, 0, :]       # iteration-0 integers
    coords[:, :, 1, :8] = -20.0                                                            # fully outside
    out = sb.encode_flow_token(maps, coords)
    for bi in (0, 15):
        ref = so.encode_flow_token(host(maps[bi * 4096:(bi + 1) * 4096]), host(coords[bi:bi + 1]))
        assert_bits_equal(host(out[bi:bi + 1].contiguous()), np.ascontiguousarray(ref), f"batch {bi}")
    # idempotence / determinism
    assert torch.equal(out, sb.encode_flow_token(maps, coords))


def test_pyramid_lookup_on_fused_levels_64(sb):
    """C3p on the levels the cost-volume kernel's fused epilogue wrote (64x64 -> 32, 16, 8): every level's
    81 taps bit-exact against the oracle's lookup of the SAME level with the centre scaled by 2^-l
    (common.py:245-248 convention); level-major channel order."""
    g = torch.Generator().manual_seed(31)
    f1, f2 = torch.randn(1, 128, 64, 64, generator=g), torch.randn(1, 128, 64, 64, generator=g)
    vol, lv = sb.corr.corr(cu(f1), cu(f2), pyramid_levels=3)
    pyr = [vol.view(4096, 1, 64, 64), lv[0], lv[1], lv[2]]
    assert [tuple(p.shape[-2:]) for p in pyr] == [(64, 64), (32, 32), (16, 16), (8, 8)]
    coords = cases.coords_grid(1, 64, 64) + torch.randn(1, 2, 64, 64, generator=g) * 3.0
    coords[0, :, 0, :] = cases.coords_grid(1, 64, 64)[0, :, 0, :]           # exact integers
    coords[0, :, 1, :4] = -40.0                                              # outside every level
    coords[0, :, 2, :4] = 63.0
    out = sb.encode_flow_token_pyramid(pyr, cu(coords))
    assert out.shape == (1, 4 * 81, 64, 64) and tuple(out.stride()) == (4 * 81 * 4096, 1, 64 * 4 * 81, 4 * 81)
    got = host(out.contiguous())
    for l, pm in enumerate(pyr):
        ref = so.encode_flow_token(host(pm), coords.numpy(), coord_scale=1.0 / (1 << l))
        assert_bits_equal(got[:, l * 81:(l + 1) * 81], np.ascontiguousarray(ref), f"pyramid level {l}")
    ref_all = so.encode_flow_token_pyramid([host(pm) for pm in pyr], coords.numpy())
    assert_bits_equal(got, np.ascontiguousarray(ref_all), "all levels")


@pytest.mark.parametrize("r", [0, 1, 2, 3, 5, 7])
def test_lookup_other_radii(sb, r):
    """decoder.py:295-315 calls encode_flow_token with radii other than 4 in non-default branches
    (r = 0 is the single centre sample): generic kernel, bit-exact against the oracle."""
    c = cases.lookup_small()
    out = sb.encode_flow_token(cu(c["cost_maps"]), cu(c["coords"]), r=r)
    side = 2 * r + 1
    assert out.shape == (2, side * side, 3, 5)
    ref = so.encode_flow_token(c["cost_maps"].numpy(), c["coords"].numpy(), r=r)
    assert_bits_equal(host(out.contiguous()), np.ascontiguousarray(ref), f"r={r}")
    g = golden("lookup_small")
    if f"out_r{r}" in g.files:
        assert_bits_equal(host(out.contiguous()), g[f"out_r{r}"], f"r={r} vs the reference's own output")
    c64 = cases.lookup_64()
    out = sb.encode_flow_token(cu(c64["cost_maps"]), cu(c64["coords"]), r=r)
    ref = so.encode_flow_token(c64["cost_maps"].numpy(), c64["coords"].numpy(), r=r)
    assert_bits_equal(host(out.contiguous()), np.ascontiguousarray(ref), f"64x64 maps, r={r}")


def test_bilinear_sampler(sb):
    g = golden("bilinear_sampler")
    out, m = sb.bilinear_sampler(cu(g["img"]), cu(g["pts"]), mask=True)
    assert_bits_equal(host(out), g["out"], "bilinear_sampler")
    assert_bits_equal(host(m), g["mask"], "bilinear_sampler mask")


# ===================================================================== W1
@pytest.mark.parametrize("name", ["warp_small", "warp_flow2"])
def test_warp_cases(sb, name):
    c = getattr(cases, name)()
    g = golden(name)
    check_inputs(g, *c.values())
    out = host(sb.warp(cu(c["x"]), cu(c["flo"])))
    assert max_abs(out, g["out"]) <= 1e-3                     # the contract
    assert_bits_equal(out, g["out"], "warp vs reference golden")   # and in fact bit-exact
    # mode='nearest' (core/warp_utils.py:76-77; no caller in the reference, part of the mirrored signature)
    near = host(sb.warp(cu(c["x"]), cu(c["flo"]), mode="nearest"))
    assert_bits_equal(near, g["out_nearest"], "nearest-mode warp vs reference golden")
    assert_bits_equal(near, so.warp(c["x"].numpy(), c["flo"].numpy(), mode="nearest"), "nearest-mode warp vs oracle")
    with pytest.raises(ValueError, match="mode"):
        sb.warp(cu(c["x"]), cu(c["flo"]), mode="bicubic")


def test_warp_512_golden(sb):
    c = cases.warp_512()
    g = golden("warp_512")
    check_inputs(g, *c.values())
    out = host(sb.warp(cu(c["x"]), cu(c["flo"])))
    sy, sx = cases.WARP_512_SAMPLE
    assert_bits_equal(out[..., sy, sx], g["out_sample"], "warp 512")


@pytest.mark.parametrize("size", [256, 1024, 2048])
def test_warp_sweep_vs_oracle(sb, size):
    """config 5 sizes; fused mask multiply and overlap output included."""
    g = torch.Generator().manual_seed(size)
    x = torch.rand(1, 6, size, size, generator=g) * 255.0
    x[:, 3:] = (x[:, 3:] > 60).float()
    flo = torch.randn(1, 2, size, size, generator=g) * 4.0
    occ = (torch.rand(1, 1, size, size, generator=g) < 0.8).float()
    out, ov = sb.warp(cu(x), cu(flo), mul_mask=cu(occ), return_overlap=True)
    ref, rov = so.warp(x.numpy(), flo.numpy(), mul_mask=occ.numpy(), return_overlap=True)
    assert_bits_equal(host(out), ref, "warp * mask")
    assert_bits_equal(host(ov), rov, "overlap mask")


def test_warp_edge_cases(sb):
    # empty batch, 1x1 image, NaN flow, ragged C
    assert sb.warp(torch.zeros(0, 6, 8, 8, device="cuda"), torch.zeros(0, 2, 8, 8, device="cuda")).shape == (0, 6, 8, 8)
    x = torch.rand(1, 5, 7, 9) * 10
    flo = torch.randn(1, 2, 7, 9)
    flo[0, 0, 3, 3] = float("nan")
    out = host(sb.warp(cu(x), cu(flo)))
    ref = so.warp(x.numpy(), flo.numpy())
    assert np.array_equal(np.isnan(out), np.isnan(ref))
    assert max_abs(np.nan_to_num(out), np.nan_to_num(ref)) == 0.0
    with pytest.raises(ValueError):
        sb.warp(torch.zeros(1, 6, 8, 8, device="cuda"), torch.zeros(1, 2, 8, 9, device="cuda"))


def test_tiled_warps_are_bit_identical_to_the_default_kernels(sb):
    """The shared-memory-staged (TMA tile) forms of the flow and homography warps — opt-in through
    sb_tune(SB_TUNE_WARP_TILED = 12, 1), measured slower than the per-pixel kernels — give the same bits: smooth and
    noisy flows (outliers leave the staged box), non-finite flows, partial tiles, every channel count."""
    lib = sb._lib.load()
    g = torch.Generator().manual_seed(21)
    cases_ = []
    for (b, c, h, w) in ((2, 6, 72, 136), (1, 3, 40, 64), (1, 2, 128, 128), (1, 1, 24, 40), (2, 6, 512, 512)):
        x = torch.rand(b, c, h, w, generator=g) * 255
        smooth = torch.nn.functional.interpolate(torch.randn(b, 2, max(h // 8, 2), max(w // 8, 2), generator=g) * 3,
                                                 size=(h, w), mode="bilinear", align_corners=True)
        noisy = torch.randn(b, 2, h, w, generator=g) * 6
        bad = smooth.clone()
        for i, v in enumerate([float("nan"), float("inf"), -float("inf"), 1e30, -1e30, 3e9]):
            bad[0, i % 2, 3 + i, 5 + 2 * i] = v
        cases_.append((cu(x), [cu(smooth), cu(noisy), cu(bad), cu(smooth + 40.0)]))
    src = torch.tensor([[0.0, 0.0], [1.0, 0.0], [0.0, 1.0], [1.0, 1.0]])
    try:
        for x, flows in cases_:
            b, c, h, w = x.shape
            mask = (torch.rand(b, 1, h, w, device="cuda") < 0.8).float()
            thetas = [torch.eye(3, device="cuda")[None].repeat(b, 1, 1) + 0.05 * torch.randn(b, 3, 3, device="cuda") * s_
                      for s_ in (0.2, 1.0, 4.0)]
            outs = {}
            for mode in (0, 1):
                assert lib.sb_tune(12, mode) == 0
                res = [sb.warp(x, f) for f in flows]
                if c == 6:
                    res += list(sb.warp(x, flows[0], mul_mask=mask, return_overlap=True))
                for th in thetas:
                    o, idx = sb.torch_homo_transform.transformer(x, th, (h + 7, w - 3), return_indices=True)
                    res += [o, idx.float()]
                    if c == 3:
                        res.append(sb.torch_homo_transform.transformer(x, th, (h, w), append_ones=3))
                outs[mode] = res
            for a_, b_ in zip(outs[0], outs[1]):
                assert_bits_equal(host(a_).view(np.uint32), host(b_).view(np.uint32), f"tiled vs per-pixel, shape {tuple(x.shape)}")
    finally:
        lib.sb_tune(12, 0)


# ===================================================================== W2
@pytest.mark.parametrize("name", ["homo_small", "homo_theta1", "homo_degenerate"])
def test_homo_cases(sb, name):
    c = getattr(cases, name)()
    g = golden(name)
    check_inputs(g, c["U"], c["theta"])
    out, idx = sb.torch_homo_transform.transformer(cu(c["U"]), cu(c["theta"]), c["out_size"], return_indices=True)
    out, idx = host(out), host(idx)
    assert_bits_equal(idx, g["idx"], "integer grid indices vs reference golden")
    if name == "homo_degenerate":
        ref = so.homo_transformer(c["U"].numpy(), c["theta"].numpy(), c["out_size"])
        fin = np.isfinite(ref) & (np.abs(ref) < 1e6)
        assert max_abs(np.where(fin, out, 0), np.where(fin, ref, 0)) <= 1e-3
    else:
        assert_bits_equal(out, g["out"], "warped values vs reference golden")


def test_homo_512_golden_and_tensor_sizes(sb):
    c = cases.homo_512()
    g = golden("homo_512")
    check_inputs(g, c["U"], c["theta"])
    # 0-dim int tensors as out_size, like flowHomoAdpater.py:292
    size = (torch.tensor(512, dtype=torch.int32), torch.tensor(512, dtype=torch.int32))
    out, idx = sb.torch_homo_transform.transformer(cu(c["U"]), cu(c["theta"]), size, return_indices=True)
    out, idx = host(out), host(idx)
    sy, sx = cases.HOMO_512_SAMPLE
    assert_bits_equal(idx[..., sy, sx], g["idx_sample"], "indices")
    assert_bits_equal(out[..., sy, sx], g["out_sample"], "values")
    assert_bits_equal(np.packbits(out[0, 3] > 0.5), g["mask_bits"], "thresholded mask")


@pytest.mark.parametrize("size", [256, 2048])
def test_homo_sweep_vs_oracle(sb, size):
    g = torch.Generator().manual_seed(size + 1)
    U = torch.rand(2, 6, size, size, generator=g) * 255.0
    U[:, 3:] = 1.0
    theta = torch.eye(3).repeat(2, 1, 1) + 0.04 * torch.randn(2, 3, 3, generator=g)
    out, idx = sb.torch_homo_transform.transformer(cu(U), cu(theta), (size + 13, size - 7), return_indices=True)
    ref, ridx = so.homo_transformer(U.numpy(), theta.numpy(), (size + 13, size - 7), return_indices=True)
    assert_bits_equal(host(idx), ridx, "indices")
    assert_bits_equal(host(out), ref, "values")


# ===================================================================== W3
def test_tps_case(sb):
    c = cases.tps_small()
    g = golden("tps_small")
    check_inputs(g, c["U"], c["source"], c["target"])
    out, idx = sb.torch_tps_transform.transformer(cu(c["U"]), cu(c["source"]), cu(c["target"]), c["out_size"],
                                                  return_indices=True)
    out, idx = host(out), host(idx)
    # The reference sums T @ basis with BLAS in an unspecified order and solves the system with a
    # different LU (GPU vs CPU): coordinates agree to ~1e-5, so samples sitting on an integer
    # boundary may floor differently; those are excluded and counted, the rest meets 1e-3.
    mism = (idx != g["idx"]).any(axis=1)
    assert mism.mean() < 0.01
    ok = ~mism[:, None].repeat(6, 1)
    assert max_abs(np.where(ok, out, 0), np.where(ok, g["out"], 0)) <= 1e-3


def test_tps_169_points_vs_oracle(sb):
    """12x12 mesh (169 control points, Homography/network.py:9-10) at 256^2."""
    g = torch.Generator().manual_seed(41)
    ys, xs = torch.meshgrid(torch.linspace(-1, 1, 13), torch.linspace(-1, 1, 13), indexing="ij")
    src = torch.stack([xs, ys], -1).reshape(1, -1, 2)
    tgt = src + 0.02 * torch.randn(1, 169, 2, generator=g)
    yy, xx = torch.meshgrid(torch.linspace(0, 3.0, 256), torch.linspace(0, 4.0, 256), indexing="ij")
    U = torch.stack([torch.sin(xx) + yy, torch.cos(yy), xx * 0.2, torch.ones_like(xx), torch.ones_like(xx), torch.ones_like(xx)], 0)[None].contiguous()
    T = sb.torch_tps_transform.solve_system(cu(src), cu(tgt))
    out, idx = sb.torch_tps_transform.transformer(cu(U), cu(src), cu(tgt), (256, 256), return_indices=True)
    # same T on both sides; the kernel's log is lg2.approx * ln2 and its sum runs in fp32 groups of 8
    # control points added in fp64 (measured: 2.3e-4 of the taps floor differently, 1.5e-5 elsewhere)
    ref, ridx = so.tps_transformer(U.numpy(), src.numpy(), tgt.numpy(), (256, 256), return_indices=True, T=host(T))
    mism = (host(idx) != ridx).any(axis=1)
    assert mism.mean() < 0.002
    ok = ~mism[:, None].repeat(6, 1)
    assert max_abs(np.where(ok, host(out), 0), np.where(ok, ref, 0)) <= 1e-3


def test_tps_log_variants_agree(sb):
    """SB_TUNE_TPS_LOG = 1 (libdevice logf) against the default lg2.approx * ln2, both TPS kernels, on
    shapes with a partial last chunk and a control-point count that is not a multiple of 8."""
    lib = sb._lib.load()
    g = torch.Generator().manual_seed(47)
    ys, xs = torch.meshgrid(torch.linspace(-1, 1, 7), torch.linspace(-1, 1, 7), indexing="ij")
    src = torch.stack([xs, ys], -1).reshape(1, -1, 2).repeat(2, 1, 1)
    tgt = src + 0.02 * torch.randn(2, 49, 2, generator=g)
    yy, xx = torch.meshgrid(torch.linspace(0, 3.0, 75), torch.linspace(0, 4.0, 91), indexing="ij")
    U = torch.stack([torch.sin(xx) + yy, torch.cos(yy), xx * 0.2], 0)[None].repeat(2, 1, 1, 1).contiguous()
    kw = 0.01 * torch.randn(2, 49, 2, generator=g)
    aw = torch.tensor([[0.01, -0.02], [1.0, 0.01], [-0.01, 1.0]]).repeat(2, 1, 1)
    got = {}
    try:
        for mode in (1, 0):
            assert lib.sb_tune(8, mode) == 0
            out, idx = sb.torch_tps_transform.transformer(cu(U), cu(src), cu(tgt), (61, 83), return_indices=True)
            ko, kg = sb.kornia_tps.warp_image_tps(cu(U), cu(src * 0.45 + 0.5), cu(kw), cu(aw), return_grid=True)
            got[mode] = [host(t) for t in (out, idx, ko, kg)]
    finally:
        lib.sb_tune(8, 0)
    mism = (got[0][1] != got[1][1]).any(axis=1)
    assert mism.mean() < 0.002
    ok = ~mism[:, None].repeat(3, 1)
    assert max_abs(np.where(ok, got[0][0], 0), np.where(ok, got[1][0], 0)) <= 1e-4
    assert max_abs(got[0][3], got[1][3]) <= 2e-6
    assert max_abs(got[0][2], got[1][2]) <= 1e-3
    ref, ridx = so.tps_transformer(U.numpy(), src.numpy(), tgt.numpy(), (61, 83), return_indices=True,
                                   T=host(sb.torch_tps_transform.solve_system(cu(src), cu(tgt))))
    for mode in (0, 1):
        mm = (got[mode][1] != ridx).any(axis=1)
        assert mm.mean() < 0.002
        okk = ~mm[:, None].repeat(3, 1)
        assert max_abs(np.where(okk, got[mode][0], 0), np.where(okk, ref, 0)) <= 1e-3


# ===================================================================== N1 (next row 1)
def test_gma_golden(sb):
    c = cases.gma_small()
    g = golden("gma_small")
    check_inputs(g, *c.values())
    attn = sb.gma.attention(cu(c["fmap"]), cu(c["w_qk"]), heads=1)
    a = host(attn)
    assert a.shape == g["attn"].shape
    # kernel check: the same contraction on the bf16-rounded projections torch produced on this GPU
    # (fp64), + TF32 rounding of the probabilities (2^-11)
    q, k = sb.gma.project_qk(cu(c["fmap"]), cu(c["w_qk"]), heads=1)
    a_bf = so.gma_attention_from_qk(host(q), host(k), bf16_inputs=True).reshape(a.shape)
    assert max_abs(a, a_bf) <= 6e-4
    assert abs(float(a.sum(-1).mean()) - 1.0) <= 1e-3
    # contract vs the fp32 reference: < 3 % of the row maximum (measured 1.7 %; the bf16 rounding of
    # q, k flips with the last bits of the library conv that produced them)
    assert (np.abs(a - g["attn"]).max(-1) / g["attn"].max(-1)).max() <= 3e-2
    out = host(sb.gma.aggregate(attn, cu(c["motion"]), cu(c["w_v"]), cu(c["gamma"])))
    scale = float(np.abs(g["out"]).max())
    # kernel check: fp64 aggregate of the kernel's own attention (TF32 operands: 2^-11 relative)
    assert max_abs(out, so.gma_aggregate(a, c["motion"].numpy(), c["w_v"].numpy(), float(c["gamma"]))) <= 2e-3 * scale
    assert max_abs(out, g["out"]) <= 1e-2 * scale
    # module mirrors with the reference's constructor arguments (decoder.py:197, gru.py:316)
    att_m = sb.gma.Attention(args=None, dim=128, heads=1, max_pos_size=160, dim_head=128).cuda()
    agg_m = sb.gma.Aggregate(args=None, dim=128, dim_head=128, heads=1).cuda()
    att_m.to_qk.weight.data.copy_(c["w_qk"]); agg_m.to_v.weight.data.copy_(c["w_v"]); agg_m.gamma.data.copy_(c["gamma"])
    with torch.no_grad():
        o2 = host(agg_m(att_m(cu(c["fmap"])), cu(c["motion"])))
    assert max_abs(o2, out) == 0.0


@pytest.mark.parametrize("h,w", [(12, 16), (30, 46), (6, 6), (64, 64)])
def test_gma_attention_two_pass_vs_unfused(sb, h, w):
    """The fused two-pass softmax (statistics pass + normalising pass over the same tcgen05 contraction)
    against the unfused path (fp32 logits -> sb_softmax_rows) on the same projections: token counts that
    are / are not multiples of the 128-wide tiles; rows sum to 1; probabilities are TF32 values."""
    gen = torch.Generator(device="cuda").manual_seed(72 + h)
    fmap = torch.randn(2, 128, h, w, device="cuda", generator=gen)
    w_qk = (torch.rand(256, 128, 1, 1, device="cuda", generator=gen) * 2 - 1) * 0.25
    n = h * w
    attn = sb.gma.attention(fmap, w_qk).view(2, n, n)
    q, k = sb.gma.project_qk(fmap, w_qk)
    sim = sb.corr.corr(q, k).view(2, n, n)
    ref64 = torch.softmax(sim.double(), dim=-1)
    unf = sb.gma.softmax_rows_(sim.clone(), to_tf32=True)
    assert float((attn.sum(-1) - 1).abs().max()) <= 2e-3
    # same logits on both sides: differences are ex2.approx vs expf (~2 ulp) and TF32 rounding flips (2^-11)
    assert float(((attn - unf).abs() / (unf + 1e-6)).max()) <= 2e-3
    assert float(((attn.double() - ref64).abs() / (ref64 + 1e-6)).max()) <= 1e-3
    assert int((attn.view(torch.int32) & 0x1fff).count_nonzero()) == 0          # low 13 mantissa bits clear


def test_gma_bf16_attention_option(sb):
    """Opt-in bf16 probabilities: aggregate within 1e-2 * scale of the fp32 reference, and the kernel
    within 1e-3 * scale of fp64 on its own bf16 operands."""
    c = cases.gma_small()
    g = golden("gma_small")
    attn = sb.gma.attention(cu(c["fmap"]), cu(c["w_qk"]), heads=1, dtype=torch.bfloat16)
    assert attn.dtype == torch.bfloat16 and tuple(attn.shape) == g["attn"].shape
    out = host(sb.gma.aggregate(attn, cu(c["motion"]), cu(c["w_v"]), cu(c["gamma"])))
    scale = float(np.abs(g["out"]).max())
    assert max_abs(out, g["out"]) <= 1e-2 * scale
    v = torch.nn.functional.conv2d(cu(c["motion"]), cu(c["w_v"])).view(2, 128, -1).bfloat16().double()
    ref = cu(c["motion"]).double() + 0.7 * torch.einsum("bij,bdj->bdi", attn.view(2, 192, 192).double(), v).view(2, 128, 12, 16)
    assert float((torch.from_numpy(out).cuda().double() - ref).abs().max()) <= 1e-3 * scale


def test_gma_full_size_vs_torch(sb):
    """512^2 shape (N = 4096 tokens), batch 2: attn @ v against torch fp32 on the same attention."""
    gen = torch.Generator(device="cuda").manual_seed(71)
    fmap = torch.randn(2, 128, 64, 64, device="cuda", generator=gen)
    w_qk = (torch.rand(256, 128, 1, 1, device="cuda", generator=gen) * 2 - 1) * 0.25
    w_v = (torch.rand(128, 128, 1, 1, device="cuda", generator=gen) * 2 - 1) * 0.09
    gamma = torch.tensor([0.5], device="cuda")
    attn = sb.gma.attention(fmap, w_qk)
    assert tuple(attn.shape) == (2, 1, 4096, 4096)
    rs = attn.sum(-1)
    assert float((rs - 1).abs().max()) <= 2e-3
    out = sb.gma.aggregate(attn, fmap, w_v, gamma)
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        v = torch.nn.functional.conv2d(fmap, w_v).view(2, 128, 4096)
        ref = fmap + gamma * torch.einsum("bij,bdj->bdi", attn.view(2, 4096, 4096).double(), v.double()).float().view(2, 128, 64, 64)
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev
    err = float((out - ref).abs().max())
    assert err <= 2e-3 * float(ref.abs().max()), err
    # ragged token count (N = 30 x 46 = 1380: not a multiple of the 128-query tile or the 32-key step)
    fm2 = torch.randn(1, 128, 30, 46, device="cuda", generator=gen)
    a2 = sb.gma.attention(fm2, w_qk)
    o2 = sb.gma.aggregate(a2, fm2, w_v, gamma)
    v2 = torch.nn.functional.conv2d(fm2, w_v).view(1, 128, 1380)
    r2 = fm2 + gamma * torch.einsum("bij,bdj->bdi", a2.view(1, 1380, 1380).double(), v2.double()).float().view(1, 128, 30, 46)
    assert float((o2 - r2).abs().max()) <= 2e-3 * float(r2.abs().max())


# ===================================================================== N3 (next row 3)
def test_gma_ragged_token_count_and_autograd_guard(sb):
    """Token maps whose size is not a multiple of 4 (the reference handles e.g. 65 x 67): the contraction runs
    on the tcgen05 kernel with a padded pitch, softmax / aggregate on library code; and a requires_grad input
    with autograd enabled is refused instead of silently detached."""
    g = torch.Generator().manual_seed(77)
    h, w = 5, 7
    fmap, motion = torch.randn(2, 128, h, w, generator=g), torch.randn(2, 128, h, w, generator=g)
    w_qk, w_v = torch.randn(256, 128, 1, 1, generator=g) * 0.1, torch.randn(128, 128, 1, 1, generator=g) * 0.1
    gamma = torch.tensor([0.7])
    attn = sb.gma.attention(cu(fmap), cu(w_qk), heads=1)
    ref_attn = so.gma_attention(fmap.numpy(), w_qk.numpy(), heads=1)
    assert attn.shape == (2, 1, h * w, h * w)
    assert max_abs(host(attn), ref_attn) <= 3e-2 * float(ref_attn.max())
    out = sb.gma.aggregate(attn, cu(motion), cu(w_v), cu(gamma))
    ref_out = so.gma_aggregate(host(attn), motion.numpy(), w_v.numpy(), gamma.numpy())
    assert max_abs(host(out), ref_out) <= 1e-2 * float(np.abs(ref_out).max())
    with torch.enable_grad():
        x = cu(fmap).requires_grad_(True)
        with pytest.raises(RuntimeError, match="inference-only"):
            sb.corr.corr(x, x)
        with pytest.raises(RuntimeError, match="inference-only"):
            sb.warp(x[:, :6], cu(torch.zeros(2, 2, h, w)))


def test_gemm_nt_tf32_vs_fp64(sb):
    gen = torch.Generator(device="cuda").manual_seed(81)
    for bh, m, n, k in ((2, 256, 128, 64), (3, 200, 136, 100), (1, 1024, 1024, 1024)):
        a = torch.randn(bh, m, k, device="cuda", generator=gen)
        b = torch.randn(bh, n, k, device="cuda", generator=gen)
        d = sb.udis2_homography.gemm_nt_tf32(a, b)
        ref = torch.einsum("bmk,bnk->bmn", a.double(), b.double())
        err = float((d.double() - ref).abs().max())
        # TF32 operands (10-bit mantissa, truncated by the tensor core): ~2^-10 * sqrt(K) * |a||b|
        assert err <= 8e-3 * (k ** 0.5), (bh, m, n, k, err)


def test_ccl_golden_and_full_size(sb):
    c = cases.ccl_small()
    g = golden("ccl_small")
    check_inputs(g, *c.values())
    flow = host(sb.udis2_homography.CCL(cu(c["feature_1"]), cu(c["feature_2"])))
    assert flow.shape == g["flow"].shape
    # TF32 contraction (10-bit mantissa): with only 64 channels every normalised element is large
    # (~1/8), which is the worst case for the x10 softmax; the network's 1024-channel shape is ~3e-4
    assert max_abs(flow, g["flow"]) <= 1e-2
    # the shape the network uses (network.py:125-130): [B, 1024, 32, 32]; also a soft softmax (scale 1),
    # which is far more sensitive to the correlation values than the reference's scale 10
    gen = torch.Generator().manual_seed(82)
    f1 = torch.relu(torch.randn(2, 1024, 32, 32, generator=gen))
    f2 = torch.roll(f1, shifts=(2, -3), dims=(2, 3)) + 0.5 * torch.relu(torch.randn(2, 1024, 32, 32, generator=gen))
    for scale, tol in ((10.0, 2e-3), (1.0, 2e-3)):
        got = host(sb.udis2_homography.CCL(cu(f1), cu(f2), softmax_scale=scale))
        ref = so.ccl(f1.numpy(), f2.numpy(), softmax_scale=scale)
        assert max_abs(got, ref) <= tol, (scale, max_abs(got, ref))


@pytest.mark.parametrize("h,w", [(9, 7), (6, 10), (36, 40)])
def test_ccl_shapes(sb, h, w):
    """H*W % 4 != 0 (row-streaming fallback kernel), W % 4 != 0 (a CTA's positions wrap image rows) and
    H*W > 1024 (the staged kernel walks q in chunks with a running softmax)."""
    gen = torch.Generator().manual_seed(83 + h)
    f1 = torch.relu(torch.randn(2, 64, h, w, generator=gen))
    f2 = torch.roll(f1, shifts=(1, -2), dims=(2, 3)) + 0.3 * torch.relu(torch.randn(2, 64, h, w, generator=gen))
    got = host(sb.udis2_homography.CCL(cu(f1), cu(f2)))
    ref = so.ccl(f1.numpy(), f2.numpy())
    assert got.shape == ref.shape == (2, 2, h, w)
    assert max_abs(got, ref) <= 1e-2 * max(1.0, max(h, w) / 12.0), max_abs(got, ref)


def test_shutdown_releases_library_state_and_library_stays_usable(sb):
    lib = sb._lib.load()
    x = torch.rand(1, 3, 16, 16, device="cuda")
    flo = torch.zeros(1, 2, 16, 16, device="cuda")
    a = sb.warp(x, flo)
    lib.sb_tune(0, 1)
    assert lib.sb_shutdown() == 0 and lib.sb_shutdown() == 0        # idempotent
    assert lib.sb_launch_count() == 0
    cm = torch.rand(64, 1, 8, 8, device="cuda")
    tok = sb.encode_flow_token(cm, sb.lookup.coords_grid(1, 8, 8, device="cuda"))   # needs the debug word again
    assert tuple(tok.shape) == (1, 81, 8, 8)
    assert torch.equal(sb.warp(x, flo), a)


def test_next_rows_empty_and_bad_inputs(sb):
    """Empty batches return empty results of the reference's shape; wrong devices / shapes raise."""
    z = lambda *sh: torch.zeros(*sh, device="cuda")
    assert tuple(sb.decoder.upsample_flow(z(0, 2, 8, 8), z(0, 576, 8, 8)).shape) == (0, 2, 64, 64)
    assert tuple(sb.udis2_homography.CCL(z(0, 64, 8, 8), z(0, 64, 8, 8)).shape) == (0, 2, 8, 8)
    assert tuple(sb.kornia_tps.warp_image_tps(z(0, 6, 8, 8), z(0, 4, 2), z(0, 4, 2), z(0, 3, 2)).shape) == (0, 6, 8, 8)
    assert tuple(sb.kornia_tps.grid_sample(z(0, 3, 8, 8), z(0, 5, 7, 2)).shape) == (0, 3, 5, 7)
    assert tuple(sb.gma.attn_matmul_v(z(0, 64, 64), z(0, 128, 64)).shape) == (0, 128, 64)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        sb.decoder.upsample_flow(torch.zeros(1, 2, 8, 8), torch.zeros(1, 576, 8, 8))
    with pytest.raises(ValueError):
        sb.decoder.upsample_flow(z(1, 2, 8, 8), z(1, 64, 8, 8))
    with pytest.raises(ValueError):
        sb.udis2_homography.CCL(z(1, 64, 8, 8), z(1, 64, 8, 9))
    with pytest.raises(RuntimeError, match="dim_head"):
        sb.gma.attn_matmul_v(z(1, 64, 64), z(1, 96, 64))
    # a key count that is not a multiple of 4 is served by the library GEMM (the reference accepts any size)
    assert tuple(sb.gma.attn_matmul_v(z(1, 66, 66), z(1, 128, 66)).shape) == (1, 128, 66)


# ===================================================================== N4 (next row 4)
def _pe_args(c):
    return [cu(c[k]) for k in ("x", "w1", "b1", "w2", "b2", "w3", "b3")]


def test_patch_embed_golden(sb):
    """N4: the fused tcgen05 conv stack vs (1) the oracle on bf16-rounded operands (what the kernel contracts;
    tight), (2) the reference's own PatchEmbed output (contract: 1e-2 of the output scale, bf16 operands vs fp32)."""
    c = cases.patch_embed_small()
    g = golden("patch_embed_small")
    check_inputs(g, *c.values())
    out = host(sb.encoder.patch_embed_proj(*_pe_args(c)))
    assert out.shape == (6, 64, 8, 8)
    scale = float(np.abs(g["proj"]).max())
    a = {k: v.numpy() for k, v in c.items()}
    assert max_abs(out, so.patch_embed_proj(**a, bf16_operands=True)) <= 2e-3 * scale      # fp32 vs fp64 accumulation, rare bf16 re-roundings
    assert max_abs(out, g["proj"]) <= 1e-2 * scale
    assert max_abs(out, so.patch_embed_proj(**a)) <= 1e-2 * scale
    # module form with the reference's parameter names
    m = sb.encoder.PatchEmbedProj().cuda()
    m.load_state_dict({f"proj.{i}.{n}": c[f"{n[0]}{k}"] for i, k in ((0, 1), (2, 2), (4, 3)) for n in ("weight", "bias")}, strict=True)
    assert torch.equal(m(cu(c["x"])), torch.from_numpy(out).cuda())


def test_patch_embed_full_size_properties(sb):
    """Two images' worth of cost maps plus a ragged tail (8195 maps: odd count, partial last wave): contract vs the
    library convolutions on a sample, order independence (a map's result does not depend on which CTA / slot / wave
    processed it), determinism, zero maps."""
    c = cases.patch_embed_small()
    g = torch.Generator(device="cuda").manual_seed(8)
    n = 2 * 4096 + 3
    x = torch.randn(n, 1, 64, 64, device="cuda", generator=g) * 16.0
    x[17] = 0.0
    args = _pe_args(c)[1:]
    pack = sb.encoder.pack_patch_embed_weights(args[0], args[2], args[4])
    out = sb.encoder.patch_embed_proj(x, *args, pack=pack)
    torch.cuda.synchronize()
    assert out.shape == (n, 64, 8, 8) and torch.isfinite(out).all()
    sel = torch.tensor([0, 1, 2, 17, 147, 148, 149, 295, 296, 4095, 4096, 8191, 8192, 8193, 8194], device="cuda")
    ref = x[sel]
    for w, b, relu in ((args[0], args[1], True), (args[2], args[3], True), (args[4], args[5], False)):
        ref = torch.nn.functional.conv2d(ref, w, b, stride=2, padding=2)
        ref = torch.relu(ref) if relu else ref
    scale = ref.abs().max().item()
    assert (out[sel] - ref).abs().max().item() <= 1e-2 * scale
    assert torch.equal(out, sb.encoder.patch_embed_proj(x, *args, pack=pack))               # deterministic
    perm = torch.randperm(n, device="cuda", generator=g)
    out_p = sb.encoder.patch_embed_proj(x[perm].contiguous(), *args, pack=pack)
    assert torch.equal(out_p, out[perm])                                                     # order independent
    one = sb.encoder.patch_embed_proj(x[4097:4098].contiguous(), *args, pack=pack)           # a single map (half-empty pair)
    assert torch.equal(one[0], out[4097])
    z = sb.encoder.patch_embed_proj(torch.zeros(1, 1, 64, 64, device="cuda"), *args, pack=pack)
    assert torch.equal(z[0], out[17])
    with pytest.raises(NotImplementedError, match="64x64"):
        sb.encoder.patch_embed_proj(torch.zeros(2, 1, 32, 32, device="cuda"), *args)
    assert sb.encoder.patch_embed_proj(torch.zeros(0, 1, 64, 64, device="cuda"), *args).shape == (0, 64, 8, 8)


def coords_grid(batch, ht, wd):                      # core/utils/utils.py:97-100 (used by the drop-in forward below)
    ys, xs = torch.meshgrid(torch.arange(ht), torch.arange(wd), indexing="ij")
    return torch.stack([xs, ys], dim=0).float()[None].repeat(batch, 1, 1, 1)


def LinearPositionEmbeddingSine(x, dim=128, NORMALIZE_FACOR=1 / 200):      # attention.py:156-161 of the reference
    fb = torch.linspace(0, dim // 4 - 1, dim // 4).to(x.device)
    return torch.cat([torch.sin(3.14 * x[..., -2:-1] * fb * NORMALIZE_FACOR), torch.cos(3.14 * x[..., -2:-1] * fb * NORMALIZE_FACOR),
                      torch.sin(3.14 * x[..., -1:] * fb * NORMALIZE_FACOR), torch.cos(3.14 * x[..., -1:] * fb * NORMALIZE_FACOR)], dim=-1)


class _PatchEmbedShell(torch.nn.Module):
    """A module with the reference PatchEmbed's attributes (encoder.py:20-58); its forward IS the drop-in body.
    The body takes coords_grid / LinearPositionEmbeddingSine from the defining module of type(self), as it does
    for the reference class — here: this test module (restated above, network code outside the hot path)."""

    def __init__(self):
        super().__init__()
        from types import SimpleNamespace
        import stitch_b200
        self.patch_size, self.dim, self.pe, self.cfg = 8, 64, "linear", SimpleNamespace(patch_embed="single", use_rpe=False)
        self.proj = stitch_b200.encoder.PatchEmbedProj().proj
        self.ffn_with_coord = torch.nn.Sequential(torch.nn.Conv2d(128, 128, 1), torch.nn.ReLU(), torch.nn.Conv2d(128, 128, 1))
        self.norm = torch.nn.LayerNorm(128)

    def forward(self, x, *masks):
        import stitch_b200
        return stitch_b200.encoder.patch_embed_forward(self, x, *masks)


def test_patch_embed_forward_dropin(sb):
    """PatchEmbed.forward with the conv stack replaced by the kernel vs the reference module's own tokens."""
    c = cases.patch_embed_small()
    g = golden("patch_embed_small")
    m = _PatchEmbedShell()
    for i, k in ((0, 1), (2, 2), (4, 3)):
        m.proj[i].weight.data.copy_(c[f"w{k}"]); m.proj[i].bias.data.copy_(c[f"b{k}"])
    for i in (0, 2):
        m.ffn_with_coord[i].weight.data.copy_(torch.from_numpy(g[f"ffn{i}_w"])); m.ffn_with_coord[i].bias.data.copy_(torch.from_numpy(g[f"ffn{i}_b"]))
    m = m.cuda().eval()
    tokens, size = m(cu(c["x"]))
    assert tuple(size) == tuple(g["size"]) == (8, 8) and tokens.shape == g["tokens"].shape == (6, 64, 128)
    assert max_abs(host(tokens), g["tokens"]) <= 3e-2          # LayerNorm output, O(1): bf16 conv stack vs fp32
    # a pre-training mask disables the fused path: the module's own layers run, as in the reference
    mask1 = torch.zeros(6, 1, 64, 64, device="cuda")
    t2, _ = m(cu(c["x"]), mask1, None, None)
    assert max_abs(host(t2), g["tokens"]) <= 1e-2              # cuDNN convolutions (TF32 allowed by default) vs the CPU run


# ===================================================================== N2 (next row 2)
def test_upsample_flow_golden(sb):
    c = cases.upsample_small()
    g = golden("upsample_small")
    check_inputs(g, *c.values())
    out = host(sb.decoder.upsample_flow(cu(c["flow"]), cu(c["mask"])))
    assert out.shape == g["out"].shape
    assert max_abs(out, g["out"]) <= 3e-5
    assert max_abs(out, so.upsample_flow(c["flow"].numpy(), c["mask"].numpy())) <= 3e-5


def test_upsample_flow_full_size_and_properties(sb):
    gen = torch.Generator().manual_seed(61)
    flow = torch.randn(2, 2, 64, 64, generator=gen) * 4.0
    mask = torch.randn(2, 576, 64, 64, generator=gen)
    out = host(sb.decoder.upsample_flow(cu(flow), cu(mask)))
    assert max_abs(out, so.upsample_flow(flow.numpy(), mask.numpy())) <= 3e-5
    # a mask that puts all weight on the centre tap (k = 4) reproduces 8 * flow, nearest-upsampled, exactly
    onehot = torch.full((2, 576, 64, 64), -1e4)
    onehot[:, 4 * 64:5 * 64] = 0.0
    out = host(sb.decoder.upsample_flow(cu(flow), cu(onehot)))
    ref = (8.0 * flow).repeat_interleave(8, dim=2).repeat_interleave(8, dim=3).numpy()
    assert_bits_equal(out, ref, "centre-tap mask == 8 * flow upsampled")
    # ragged width (not a multiple of the 32-pixel CTA tile) and a single row
    flow = torch.randn(1, 2, 1, 45, generator=gen)
    mask = torch.randn(1, 576, 1, 45, generator=gen)
    out = host(sb.decoder.upsample_flow(cu(flow), cu(mask)))
    assert max_abs(out, so.upsample_flow(flow.numpy(), mask.numpy())) <= 3e-5


# ===================================================================== W3k
def test_tps_kornia_golden(sb):
    c = cases.tps_kornia_small()
    g = golden("tps_kornia")
    check_inputs(g, *c.values())
    kt = sb.kornia_tps
    img = cu(c["image"])
    for ac in (False, True):
        # the sampler alone, on the reference's own grid: bit-exact
        assert_bits_equal(host(kt.grid_sample(img, cu(g["grid"]), align_corners=ac)), g[f"out_ac{int(ac)}"],
                          f"grid_sample align_corners={ac}")
        out, grid = kt.warp_image_tps(img, cu(c["points_src"]), cu(g["kernel_weights"]), cu(g["affine_weights"]),
                                      align_corners=ac, return_grid=True)
        assert max_abs(host(grid), g["grid"]) <= 1e-5                       # K-term sum: fp32 groups of 8 added in fp64 (measured 4.3e-6)
        assert_bits_equal(host(out), host(kt.grid_sample(img, grid, align_corners=ac)), "fused == grid + sample")
        assert_bits_equal(host(out), so.grid_sample(c["image"].numpy(), host(grid), ac), "sampler vs oracle")
        assert max_abs(host(out), g[f"out_ac{int(ac)}"]) <= 1e-3            # the stated contract
    # the solve (pseudo-inverse, GPU vs CPU LAPACK) agrees to the conditioning of the system
    kw, aw = kt.get_tps_transform(cu(c["points_dst"]), cu(c["points_src"]))
    out = kt.warp_image_tps(img, cu(c["points_src"]), kw, aw)
    assert max_abs(host(out), g["out_ac0"]) <= 5e-2


def test_tps_kornia_169_points_vs_oracle(sb):
    g = torch.Generator().manual_seed(46)
    ys, xs = torch.meshgrid(torch.linspace(0.02, 0.98, 13), torch.linspace(0.02, 0.98, 13), indexing="ij")
    src = torch.stack([xs, ys], -1).reshape(1, -1, 2).repeat(2, 1, 1)
    kw = 0.01 * torch.randn(2, 169, 2, generator=g)
    aw = torch.tensor([[0.01, -0.02], [1.0, 0.01], [-0.01, 1.0]]).repeat(2, 1, 1) + 0.01 * torch.randn(2, 3, 2, generator=g)
    img = torch.rand(2, 6, 200, 264, generator=g) * 255.0
    out, grid = sb.kornia_tps.warp_image_tps(cu(img), cu(src), cu(kw), cu(aw), return_grid=True)
    rgrid = so.tps_kornia_grid(src.numpy(), kw.numpy(), aw.numpy(), 200, 264)
    assert max_abs(host(grid), rgrid) <= 1e-5
    assert_bits_equal(host(out), so.grid_sample(img.numpy(), host(grid), False), "sampler vs oracle on the kernel's grid")


# ===================================================================== W4
def _occ_compare(got, ref_raw, what):
    """Thresholded masks are bit-exact except where the reference's own fp32 value sits
    within 1e-5 of the 0.5 threshold (the reference's GPU scatter_add_ is order-dependent
    there); those pixels are counted and must be rare."""
    near = np.abs(ref_raw - 0.5) <= 1e-5
    bad = ((got >= 0.5) != (ref_raw >= 0.5)) & ~near
    assert int(bad.sum()) == 0, f"{what}: {int(bad.sum())} mask pixels differ"
    assert near.mean() < 1e-3


@pytest.mark.parametrize("name", ["range_small", "range_smooth"])
def test_range_map_cases(sb, name):
    c = getattr(cases, name)()
    g = golden(name)
    check_inputs(g, *c.values())
    fij, fji = cu(c["flow_ij"]), cu(c["flow_ji"])
    rm = host(sb.compute_range_map(fji))
    # fixed-point accumulation vs the reference's sequential fp32 sum: a few ulps of the sum
    assert max_abs(rm, g["range_map"]) <= 2e-6 * max(1.0, float(g["range_map"].max()))
    occ = host(sb.compute_occlusion(fij, fji, "wang", occlusion_are_zeros=True, boundaries_occluded=True))
    assert max_abs(occ, g["occ"]) <= 2e-6
    _occ_compare(occ, g["occ"], "occlusion mask")
    occ_t = host(sb.compute_occlusion(fij, fji, "wang", occlusion_are_zeros=True, threshold=True))
    _occ_compare(occ_t, g["occ"], "fused thresholded occlusion")
    assert set(np.unique(occ_t)) <= {0.0, 1.0}
    assert max_abs(host(sb.compute_occlusion(fij, fji, "wang")), g["occ_nz"]) <= 2e-6
    assert max_abs(host(sb.compute_occlusion(fij, fji, "wang", occlusion_are_zeros=True, boundaries_occluded=False)), g["occ_nb"]) <= 2e-6
    assert_bits_equal(host(sb.compute_occlusion(fij, fji, "brox")), g["brox"], "brox")
    # deterministic (the reference's GPU scatter_add_ is not)
    assert torch.equal(sb.compute_range_map(fji), sb.compute_range_map(fji))


def test_range_map_full_size_and_fanin(sb):
    g = torch.Generator().manual_seed(52)
    lo = torch.randn(16, 2, 64, 64, generator=g) * 2.0
    fl = torch.nn.functional.interpolate(lo, size=(512, 512), mode="bilinear", align_corners=True)
    rm = host(sb.compute_range_map(cu(fl)))
    ref = so.compute_range_map(fl.numpy())
    assert max_abs(rm, ref) <= 4e-6 * max(1.0, float(ref.max()))
    # extreme fan-in: every pixel of a 256^2 image lands on one target (sum = 65536 exactly)
    ys, xs = torch.meshgrid(torch.arange(256.0), torch.arange(256.0), indexing="ij")
    flow = torch.stack([100.0 - xs, 50.0 - ys], 0)[None]
    rm = host(sb.compute_range_map(cu(flow)))
    assert rm[0, 0, 50, 100] == 65536.0 and rm.sum() == 65536.0


# ===================================================================== W5
@pytest.mark.parametrize("name", ["morph_small", "morph_big"])
def test_morph_cases(sb, name):
    c = getattr(cases, name)()
    g = golden(name)
    check_inputs(g, c["mask"])
    shape = tuple(g["shape"])
    out = host(sb.preprocess_occlusion_mask(cu(c["mask"])))
    assert_bits_equal(out > 0.5, unpack_bits(g["out_bits"], shape), "19x19 open vs reference golden")
    assert set(np.unique(out)) <= {0.0, 1.0}
    out7 = host(sb.preprocess_occlusion_mask(cu(c["mask"]), kernel_size=(7, 11)))
    assert_bits_equal(out7 > 0.5, unpack_bits(g["out7_bits"], shape), "7x11 open")


@pytest.mark.parametrize("shape,k", [((16, 1, 512, 512), (19, 19)), ((1, 1, 527, 555), (19, 19)), ((2, 1, 100, 1000), (33, 5)),
                                     ((1, 1, 5, 7), (19, 19)), ((1, 1, 300, 300), (1, 1))])
def test_morph_vs_oracle(sb, shape, k):
    g = torch.Generator().manual_seed(shape[2] + k[0])
    lo = torch.rand(shape[0], 1, max(shape[2] // 16, 2), max(shape[3] // 16, 2), generator=g)
    m = torch.nn.functional.interpolate(lo, size=shape[2:], mode="bilinear", align_corners=False) * 1.3
    for bz in (True, False):
        out = host(sb.composition.morph_open(cu(m), k, border_is_zero=bz))
        assert_bits_equal(out, so.morph_open(m.numpy(), k, border_is_zero=bz), f"open {k} border_is_zero={bz}")
    # idempotence of a morphological open
    once = sb.composition.morph_open(cu(m), k)
    assert torch.equal(once, sb.composition.morph_open(once, k))


# ===================================================================== W6 / W7 / W8
def _composite_inputs(seed, h, w):
    g = torch.Generator().manual_seed(seed)
    def sixpack():
        t = torch.rand(1, 6, h, w, generator=g) * 255.0
        mask = (torch.rand(1, 1, h, w, generator=g) > 0.3).float() + 1e-7 * torch.randn(1, 1, h, w, generator=g)
        t[:, 3:] = mask
        t[:, :3] *= (mask > 0.5)
        return t
    return sixpack(), sixpack(), sixpack(), (torch.rand(1, 1, h, w, generator=g) > 0.2).float()


@pytest.mark.parametrize("hw", [(64, 80), (527, 555)])
@pytest.mark.parametrize("with_occ", [True, False])
def test_composite_vs_oracle(sb, hw, with_occ):
    h1, h2, fw, occ = _composite_inputs(hw[0], *hw)
    r = sb.composite_test_out(cu(h1), cu(h2), cu(fw), cu(occ) if with_occ else None)
    ref = so.composite_test_out(h1.numpy(), h2.numpy(), fw.numpy(), occ.numpy() if with_occ else None)
    for k in ("final_warp_output", "output2", "mask1", "mask2"):
        assert_bits_equal(host(r[k]), ref[k], k)
    assert r["blend_image"].dtype == torch.uint8
    # blend: 0/0 -> NaN -> 0 on both sides; compare everywhere (same fp32 op order)
    assert_bits_equal(host(r["blend_image"]), ref["blend_image"], "blend_image")


def test_composite_w6_reference_block(sb):
    """W6 on the GPU against the reference's own compositing block (inputs + outputs captured inside its
    test_out_forward, tests/golden/composite_w6.npz): bit-exact values, masks and uint8 blend."""
    g = golden("composite_w6")
    # :317 as the product path does it — fused into the flow warp's epilogue (mul_mask); here the multiply alone
    fw_in = cu(g["warp_out"]) * cu(g["flow_mask"])
    assert_bits_equal(host(fw_in).view(np.uint32), g["final_warp_in"].view(np.uint32), "final_warp * flow_mask")
    r = sb.composite_test_out(cu(g["homo_output"]), cu(g["homo_output2"]), fw_in, cu(g["occlusion_mask"]))
    assert_bits_equal(host(r["final_warp_output"][:, 0:3].contiguous()).view(np.uint32), g["out_final_warp"].view(np.uint32),
                      "final_warp")
    for k in ("output1", "output2", "mask1", "mask2"):
        a, b = host(r[k].contiguous()), g["out_" + k]
        assert_bits_equal(np.isnan(a), np.isnan(b), k + " NaN pattern")
        assert_bits_equal(a.view(np.uint32)[~np.isnan(a)], b.view(np.uint32)[~np.isnan(b)], k)
    assert r["blend_image"].dtype == torch.uint8
    assert_bits_equal(host(r["blend_image"]), g["out_blend_image"], "blend_image (uint8)")
    assert_bits_equal(host(r["mask1"]) > 0.5, g["out_mask1"] > 0.5, "mask1 thresholded")
    assert_bits_equal(host(r["mask2"]) > 0.5, g["out_mask2"] > 0.5, "mask2 thresholded")


def test_build_model_golden(sb):
    c = cases.build_model_small()
    g = golden("build_model")
    check_inputs(g, *c.values())
    net_out = cu(c["net_out"])
    r = sb.build_model(lambda *a: net_out, cu(c["warp1"]), cu(c["warp2"]), cu(c["mask1"]), cu(c["mask2"]))
    for k in ("learned_mask1", "learned_mask2", "stitched_image"):
        assert_bits_equal(host(r[k]), g[k], k)


def test_tps_mix_golden(sb):
    """W8 against the reference's own tps_H_warp (tps_pipline.py:138-170 executed as written)."""
    c = cases.tps_mix_small()
    g = golden("tps_mix")
    check_inputs(g, *c.values())
    tm = sb.composition.tps_warp_mask(cu(c["tps_mask3"]))
    assert_bits_equal(host(tm), g["tps_mask"], "11x11 cv2-style open of the inverse mask")
    out2, mask2, blend = sb.composition.tps_mix_blend(cu(c["final_warp"]), cu(c["tps_warp_raw"]) * tm, tm,
                                                      cu(c["output1"]), cu(c["mask1"]))
    assert_bits_equal(host(out2), g["output2"], "output2")
    assert_bits_equal(host(mask2), g["mask2"], "mask2")
    assert_bits_equal(host(blend), g["blend"], "blend")


# ===================================================================== adapter (orchestration)
def _band_ok(shape, band):
    ok = np.zeros(shape[-2:], bool)
    ok[band:-band, band:-band] = True
    return ok


def test_adapter_train_eval_golden(sb):
    c = cases.adapter_train_eval()
    g = golden("adapter_train_eval")
    check_inputs(g, c["image1"], c["image2"], c["offsets"], *c["flows"])
    ad = sb.FlowHomoAdpater(cases.StubHomo(c["offsets"]).cuda(), cases.StubFlow([f.cuda() for f in c["flows"]]), cases.adapter_cfg())
    ad.eval()
    od = ad(cu(c["image1"]), cu(c["image2"]), type="test_eval")
    assert max_abs(host(od["H"]), g["H"]) <= 1e-4
    # The 3x3 geometry (DLT, inverse) runs on the GPU here and on the CPU in the golden run, so
    # theta differs in the last ulps and the sampler's hard-zero border may flip a border pixel.
    # Images are U(0,255) noise: a 1e-5 px coordinate wobble moves values by ~3e-3, so the bulk
    # is compared at 2e-2 max-abs and pixels off by more (border flips) are counted (< 0.5 %).
    for k in ("output_H", "output_H_inv", "final_warp_output"):
        d = np.abs(host(od[k]).astype(np.float64) - g[k])
        assert (d > 2e-2).mean() < 5e-3, (k, (d > 2e-2).mean())
    assert (host(od["overlap"]) != g["overlap"]).mean() < 5e-3
    assert (host(od["origin_occlusion_mask"]) != g["origin_occlusion_mask"]).mean() < 1e-3
    assert od["final_warp_output"].shape == (2, 6, 128, 128) and od["overlap"].shape == (2, 128, 128)


def test_adapter_test_out_golden(sb):
    c = cases.adapter_test_out()
    g = golden("adapter_test_out")
    check_inputs(g, c["image1"], c["image2"], c["offsets"], *c["flows"])
    ad = sb.FlowHomoAdpater(cases.StubHomo(c["offsets"]).cuda(), cases.StubFlow([f.cuda() for f in c["flows"]]), cases.adapter_cfg())
    ad.eval()
    od = ad(cu(c["image1"]), cu(c["image2"]), type="test_out")
    canvas = [od["width_min"], od["height_min"], od["out_height"], od["out_width"]]
    assert canvas == [int(v) for v in g["canvas"]], (canvas, g["canvas"])
    assert max_abs(host(od["I_mat"]), g["I_mat"]) <= 1e-5
    sy, sx = cases.ADAPTER_SAMPLE
    # smooth images (gradient <~ 1/px): bulk within 1e-2 (GPU-vs-CPU 3x3 inverses move the
    # coordinates by ~1e-5 px .. 1e-4 px), border flips counted
    for k in ("H_warp", "final_warp", "output1", "output2", "mask1", "mask2", "H_warp_mask"):
        got = host(od[k])[..., sy, sx]
        d = np.abs(got.astype(np.float64) - g[k + "_sample"])
        assert (d > 1e-2).mean() < 5e-3, (k, (d > 1e-2).mean(), d.max())
    db = np.abs(host(od["blend_image"])[..., sy, sx].astype(np.int32) - g["blend_image_sample"].astype(np.int32))
    assert (db > 1).mean() < 5e-3
    for k in ("occlusion_mask", "origin_occlusion_mask", "warp_input2_mask"):
        want = unpack_bits(g[k + "_bits"], tuple(g[k + "_shape"]))
        got = host(od[k]) > 0.5
        assert got.shape == want.shape
        assert (got != want).mean() < 2e-3, (k, (got != want).mean())
    assert set(np.unique(host(od["occlusion_mask"]))) <= {0.0, 1.0}


# ===================================================================== whole step
def test_hot_path_step_small(sb):
    """One full step (2 volumes + pyramid, 24 lookups, 2 homography warps, occlusion, flow warp)
    at a reduced size, every output against the oracle."""
    from stitch_b200.pipeline import HotPath, make_pair_batch
    pb = make_pair_batch(0, 2, size=128, iters=2)
    hp = HotPath(size=128, iters=2, pyramid=False)
    out = hp.step(pb.map(lambda t: t.cuda()))
    f1, f2 = pb.fmap1.numpy(), pb.fmap2.numpy()
    vol = so.corr(f1, f2)
    assert max_abs(host(out["cost_volume"]), vol) <= 1e-2 * float(np.abs(vol).max())
    maps = host(out["cost_volume"]).reshape(-1, 1, 16, 16)
    assert_bits_equal(host(out["cost_tokens"][1].contiguous()),
                      np.ascontiguousarray(so.encode_flow_token(maps, pb.coords[1, 0].numpy())), "lookup")
    occ = so.compute_occlusion_wang(pb.flow_ji.numpy(), True, threshold=True)
    assert (host(out["origin_occlusion_mask"]) != occ).mean() < 1e-4
    ref_fw, ref_ov = so.warp(host(out["output_H"]), pb.flow_ij.numpy(), mul_mask=host(out["origin_occlusion_mask"]),
                             return_overlap=True)
    assert_bits_equal(host(out["final_warp_output"]), ref_fw, "final_warp_output")
    assert_bits_equal(host(out["overlap"]), ref_ov, "overlap")


def test_config3_batch64_volume_and_lookup_indexing(sb):
    """BASELINE config 3 (batch 64 per GPU): a 4 GiB volume per direction — byte offsets beyond 2^31 and 2^32,
    262144 cost maps.  Cost volume (+ fused pyramid) and lookup are checked on batches either side of the
    2 GiB / 4 GiB boundaries; the full step then runs at batch 64."""
    b = 64
    g = torch.Generator(device="cuda").manual_seed(64)
    f1 = torch.randn(b, 256, 64, 64, device="cuda", generator=g)
    f2 = torch.randn(b, 256, 64, 64, device="cuda", generator=g)
    vol, lv = sb.corr.corr(f1, f2, pyramid_levels=3)
    assert vol.numel() * 4 == 4 << 30
    v = vol.view(b, 4096, 4096)
    for bi in (0, 31, 32, 33, 63):
        ref = torch.bmm(f1[bi:bi + 1].bfloat16().float().view(1, 256, 4096).transpose(1, 2),
                        f2[bi:bi + 1].bfloat16().float().view(1, 256, 4096))[0]
        scale = ref.abs().max().item()
        assert (v[bi] - ref).abs().max().item() <= 5e-5 * scale, bi
        # fused pyramid of the same batch == chained pooling of the volume the kernel wrote
        cm = v[bi].view(4096, 1, 64, 64)
        l1 = torch.nn.functional.avg_pool2d(cm, 2, stride=2)
        assert (lv[0].view(b, 4096, 1, 32, 32)[bi] - l1).abs().max().item() <= 1e-5 * scale, bi
        l3 = torch.nn.functional.avg_pool2d(torch.nn.functional.avg_pool2d(l1, 2, stride=2), 2, stride=2)
        assert (lv[2].view(b, 4096, 1, 8, 8)[bi] - l3).abs().max().item() <= 1e-5 * scale, bi
        del ref, l1, l3
    coords = sb.lookup.coords_grid(b, 64, 64, device="cuda") + torch.randn(b, 2, 64, 64, device="cuda", generator=g) * 3.0
    maps = vol.view(b * 4096, 1, 64, 64)
    out = sb.encode_flow_token(maps, coords)
    assert out.shape == (b, 81, 64, 64)
    for bi in (0, 31, 32, 63):
        ref = so.encode_flow_token(host(maps[bi * 4096:(bi + 1) * 4096]), host(coords[bi:bi + 1]))
        assert_bits_equal(host(out[bi:bi + 1].contiguous()), np.ascontiguousarray(ref), f"lookup, batch {bi}")
    del vol, lv, out, maps, v
    torch.cuda.empty_cache()
    # the whole step at batch 64 equals four steps at batch 16 on the same pairs (pairs are independent)
    from stitch_b200.pipeline import HotPath, make_pair_batch
    pb = make_pair_batch(0, b, size=512, iters=1).map(lambda t: t.cuda())
    hp = HotPath(size=512, iters=1, pyramid=True)
    big = hp.step(pb)
    torch.cuda.synchronize()
    for s0 in (0, 48):
        part = pb.map(lambda t: t[s0:s0 + 16].contiguous() if t.shape[0] == b else t[:, :, s0:s0 + 16].contiguous())
        small = hp.step(part)
        for k in ("final_warp_output", "overlap", "origin_occlusion_mask", "output_H", "cost_volume"):
            assert torch.equal(big[k][s0:s0 + 16], small[k]), (k, s0)
        assert torch.equal(big["cost_tokens"][0][s0:s0 + 16], small["cost_tokens"][0])
        assert torch.equal(big["cost_pyramid_back"][2].view(b, -1)[s0:s0 + 16], small["cost_pyramid_back"][2].view(16, -1))
        del small


# ===================================================================== G1 fused geometry / ones fusion
def test_homo_append_ones_equals_cat(sb):
    c = cases.homo_small()
    U3 = cu(c["U"][:, :3].contiguous())
    full, idx = sb.torch_homo_transform.transformer(torch.cat((U3, torch.ones_like(U3)), 1), cu(c["theta"]), c["out_size"],
                                                    return_indices=True)
    fused, idx2 = sb.torch_homo_transform.transformer(U3, cu(c["theta"]), c["out_size"], return_indices=True, append_ones=3)
    assert torch.equal(full, fused) and torch.equal(idx, idx2)
    g = golden("homo_small")     # the golden case IS cat(image, ones)
    assert_bits_equal(host(fused), g["out"], "append_ones vs reference golden")
    f2 = sb.torch_homo_transform.transformer(cu(c["U"][:, :2].contiguous()), cu(c["theta"]), c["out_size"], append_ones=1)
    assert torch.equal(f2, full[:, [0, 1, 3]])


def test_dlt_theta_vs_oracle(sb):
    g = golden("geometry")
    src = torch.tensor([[0.0, 0.0], [64, 0.0], [0.0, 48], [64, 48]]).unsqueeze(0).repeat(3, 1, 1)
    dst = torch.from_numpy(g["dst"])
    H = host(sb.torch_DLT.tensor_DLT(cu(src), cu(dst)))
    assert max_abs(H, g["H"]) <= 1e-4                      # reference: fp32 LU; here fp64 elimination, one rounding
    M = sb.torch_DLT.norm_matrix(64, 48)
    Hh, th, thi = sb.torch_DLT.dlt_thetas(cu(src), cu(dst), left=sb.torch_DLT._inv3(M), right=M)
    Mn = np.array(M, np.float64).reshape(3, 3)
    want = np.linalg.inv(Mn) @ g["H"].astype(np.float64) @ Mn
    want_i = np.linalg.inv(Mn) @ np.linalg.inv(g["H"].astype(np.float64)) @ Mn
    assert max_abs(host(th), want) <= 1e-4 * np.abs(want).max()
    assert max_abs(host(thi), want_i) <= 1e-4 * np.abs(want_i).max()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        sb.torch_DLT.tensor_DLT(src, dst)


def test_step_graph_replay_matches_eager(sb):
    """The whole step is sync-free and capturable; a replay reproduces the eager outputs bit-for-bit."""
    from stitch_b200.pipeline import HotPath, make_pair_batch
    pb = make_pair_batch(0, 2, size=128, iters=2).map(lambda t: t.cuda())
    hp = HotPath(size=128, iters=2, pyramid=True)
    eager = hp.step(pb)
    keep = {k: eager[k].clone() for k in ("final_warp_output", "overlap", "origin_occlusion_mask", "cost_volume", "output_H_inv")}
    tok = eager["cost_tokens"][3].clone()
    hp.capture(pb)
    out = hp.replay()
    torch.cuda.synchronize()
    for k, v in keep.items():
        assert torch.equal(out[k], v), k
    assert torch.equal(out["cost_tokens"][3], tok)


# ===================================================================== config 4: 1024 x 1024 pairs
def test_corr_1024_pair_and_pyramid(sb):
    """High-res pair: N = 16384 tokens, 1 GiB volume; W2 = 128: a 128-column tile is one target row, the
    fused pyramid epilogue drains tile pairs (POOL = 2). Levels must be bit-identical to chained avg_pool2d."""
    g = torch.Generator(device="cuda").manual_seed(4)
    f1 = torch.randn(1, 256, 128, 128, device="cuda", generator=g)
    f2 = torch.randn(1, 256, 128, 128, device="cuda", generator=g)
    vol, lv = sb.corr.corr(f1, f2, pyramid_levels=3)
    assert vol.shape == (1, 1, 128, 128, 128, 128)
    v = vol.view(16384, 16384)
    a = f1.bfloat16().float().view(256, 16384)
    b = f2.bfloat16().float().view(256, 16384)
    scale = None
    for r0 in (0, 5000, 16384 - 2048):          # row slabs: never materialise a second 1 GiB tensor
        ref = a[:, r0:r0 + 2048].t() @ b
        scale = ref.abs().max().item() if scale is None else scale
        assert (v[r0:r0 + 2048] - ref).abs().max().item() <= 5e-5 * scale
        ref32 = f1.view(256, 16384)[:, r0:r0 + 2048].t() @ f2.view(256, 16384)
        assert (v[r0:r0 + 2048] - ref32).abs().max().item() <= 1e-2 * ref32.abs().max().item()
    cm = vol.view(16384, 1, 128, 128)
    want = cm
    for l in range(3):
        want = torch.nn.functional.avg_pool2d(want, 2, stride=2)
        assert lv[l].shape == want.shape
        assert torch.equal(lv[l], want), f"level {l + 1}: max diff {(lv[l] - want).abs().max().item()}"
    # lookup on the 128x128 maps (queries of one image row), bit-exact vs the oracle
    coords = sb.lookup.coords_grid(1, 128, 128, device="cuda") + torch.randn(1, 2, 128, 128, device="cuda", generator=g) * 3
    out = sb.encode_flow_token(cm, coords)
    sel = slice(128 * 37, 128 * 38)
    ref = so.encode_flow_token(host(cm[sel]), host(coords[:, :, 37:38, :]))
    assert_bits_equal(host(out[:, :, 37:38, :].contiguous()), np.ascontiguousarray(ref), "1024^2 lookup")
