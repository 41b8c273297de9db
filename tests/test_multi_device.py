"""One process, several GPUs — the reference's own multi-GPU mechanism is nn.DataParallel (out.py:80,
evaluate.py:119): one host thread per device inside ONE process.  The library's per-device state (shared-memory
opt-ins of the big kernels, the mapped debug word) must therefore work from any thread on any device.
Skipped on a single-GPU box; run with `gpurun --gpus 2 -- python -m pytest tests/test_multi_device.py -m gpu`."""
import threading

import numpy as np
import pytest
import torch

import cases
import stitch_oracle as so
from helpers import assert_bits_equal, max_abs

pytestmark = pytest.mark.gpu


def _need_two():
    if not torch.cuda.is_available():
        pytest.fail("gpu tests need a CUDA device")
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs in one process")


def _kernels_with_large_smem(sb, dev, seed, out):
    """corr (224 KB smem), r=4 lookup, GMA aggregate and CCL — the four kernels that opt in to > 48 KB of
    dynamic shared memory — on device `dev`, each against a reference computed on the same device."""
    try:
        with torch.cuda.device(dev):
            g = torch.Generator(device=f"cuda:{dev}").manual_seed(seed)
            rnd = lambda *s: torch.randn(*s, device=f"cuda:{dev}", generator=g)
            f1, f2 = rnd(2, 256, 32, 32), rnd(2, 256, 32, 32)
            vol = sb.corr.corr(f1, f2)
            ref = torch.bmm(f1.bfloat16().float().view(2, 256, 1024).transpose(1, 2), f2.bfloat16().float().view(2, 256, 1024))
            err_corr = (vol.view(2, 1024, 1024) - ref).abs().max().item() / ref.abs().max().item()
            maps = rnd(2 * 64, 1, 64, 64)
            coords = sb.lookup.coords_grid(2, 8, 8, device=f"cuda:{dev}") * 7 + rnd(2, 2, 8, 8) * 2
            tok = sb.encode_flow_token(maps, coords)
            ref_tok = so.encode_flow_token(maps.cpu().numpy(), coords.cpu().numpy())
            attn = torch.softmax(rnd(1, 1, 1024, 1024), -1)
            fmap, w_v = rnd(1, 128, 32, 32), rnd(128, 128, 1, 1) * 0.1
            gamma = torch.full((1,), 0.5, device=f"cuda:{dev}")
            agg = sb.gma.aggregate(attn, fmap, w_v, gamma)
            v = torch.nn.functional.conv2d(fmap, w_v).view(1, 128, 1024)
            ref_agg = fmap + 0.5 * torch.bmm(attn[0], v.transpose(1, 2)).transpose(1, 2).reshape(1, 128, 32, 32)
            err_agg = (agg - ref_agg).abs().max().item() / ref_agg.abs().max().item()
            c = cases.ccl_small()
            flow = sb.udis2_homography.CCL(c["feature_1"].to(f"cuda:{dev}"), c["feature_2"].to(f"cuda:{dev}"))
            ref_flow = so.ccl(c["feature_1"].numpy(), c["feature_2"].numpy())
            torch.cuda.synchronize(dev)
            out[dev] = dict(err_corr=err_corr, tok=tok.contiguous().cpu().numpy(), ref_tok=np.ascontiguousarray(ref_tok),
                            err_agg=err_agg, flow=flow.cpu().numpy(), ref_flow=ref_flow, device=str(vol.device))
    except Exception as e:  # surfaced by the main thread
        out[dev] = e


def test_large_smem_kernels_from_two_threads_on_two_devices():
    _need_two()
    import stitch_b200 as sb
    out = {}
    # device 1 FIRST, then both concurrently, then device 0 alone again
    _kernels_with_large_smem(sb, 1, 5, out)
    assert not isinstance(out[1], Exception), out[1]
    th = [threading.Thread(target=_kernels_with_large_smem, args=(sb, d, 7 + d, out)) for d in (0, 1)]
    [t.start() for t in th]
    [t.join() for t in th]
    for d in (0, 1):
        r = out[d]
        assert not isinstance(r, Exception), (d, r)
        assert r["device"] == f"cuda:{d}"
        assert r["err_corr"] <= 5e-5, (d, r["err_corr"])
        assert_bits_equal(r["tok"], r["ref_tok"], f"lookup on cuda:{d}")
        assert r["err_agg"] <= 1e-2, (d, r["err_agg"])
        assert max_abs(r["flow"], r["ref_flow"]) <= 1e-2
    assert sb._lib.load().sb_debug_word() == 0


class _Passthrough(torch.nn.Module):
    """Stub network that is a function of its inputs only, so DataParallel replicas agree with one device."""

    def __init__(self, kind):
        super().__init__()
        self.kind = kind
        self.p = torch.nn.Parameter(torch.zeros(1))

    def forward(self, a, b, out_dict=None):
        if self.kind == "homo":
            # elementwise only: a reduction's summation order may depend on the batch size a replica sees
            return torch.tanh(a[:, 0, 5, 3:11] - b[:, 1, 7, 2:10]) * 6.0, None
        lo = torch.nn.functional.avg_pool2d(a[:, :2] - b[:, :2], 8) * 0.02
        return [torch.nn.functional.interpolate(lo, size=a.shape[-2:], mode="bilinear", align_corners=True)]


def test_adapter_under_dataparallel_matches_single_device():
    """evaluate.py:119 wraps the model in nn.DataParallel; the package's adapter must give, on two GPUs of one
    process, exactly what it gives on one."""
    _need_two()
    import stitch_b200 as sb
    g = torch.Generator().manual_seed(3)
    im1 = torch.rand(4, 3, 128, 128, generator=g) * 255
    im2 = torch.rand(4, 3, 128, 128, generator=g) * 255
    ad = sb.FlowHomoAdpater(_Passthrough("homo"), _Passthrough("flow"), cases.adapter_cfg()).cuda(0).eval()
    with torch.no_grad():
        single = ad(im1.cuda(0), im2.cuda(0), type="test_eval")
        dp = torch.nn.DataParallel(ad, device_ids=[0, 1])
        multi = dp(im1.cuda(0), im2.cuda(0), type="test_eval")
    for k in ("output_H", "output_H_inv", "final_warp_output", "overlap", "H"):
        assert multi[k].shape == single[k].shape, k
        assert torch.equal(multi[k].cpu(), single[k].cpu()), k
