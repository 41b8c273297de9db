"""Checkpoint compatibility of the module mirrors: the reference loads its checkpoints with strict=True
(evaluate.py:123, out.py:75,85), so parameter / buffer names and shapes must match exactly."""
import os
import sys

import pytest
import torch

REF = os.environ.get("STITCH_REFERENCE", "/root/reference")


def _expected_attention_state(dim, heads, max_pos_size, dim_head):
    return {"to_qk.weight": (heads * dim_head * 2, dim, 1, 1),
            "pos_emb.rel_height.weight": (2 * max_pos_size - 1, dim_head),
            "pos_emb.rel_width.weight": (2 * max_pos_size - 1, dim_head),
            "pos_emb.rel_ind": (max_pos_size, max_pos_size)}


def test_gma_modules_have_the_reference_state_dict_layout():
    import stitch_b200
    att = stitch_b200.gma.Attention(args=None, dim=128, heads=1, max_pos_size=160, dim_head=128)
    want = _expected_attention_state(128, 1, 160, 128)          # gma.py:6-18,35-52 at decoder.py:197's arguments
    got = {k: tuple(v.shape) for k, v in att.state_dict().items()}
    assert got == want
    # strict load of a reference-shaped checkpoint fragment
    sd = {k: torch.zeros(s, dtype=torch.long if k.endswith("rel_ind") else torch.float32) for k, s in want.items()}
    att.load_state_dict(sd, strict=True)
    agg = stitch_b200.gma.Aggregate(args=None, dim=128, dim_head=128, heads=1)
    assert {k: tuple(v.shape) for k, v in agg.state_dict().items()} == {"to_v.weight": (128, 128, 1, 1), "gamma": (1,)}
    agg4 = stitch_b200.gma.Aggregate(args=None, dim=128, dim_head=128, heads=4)
    assert set(agg4.state_dict()) == {"to_v.weight", "gamma", "project.weight"}


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "core")), reason="reference tree not present (GPU box)")
def test_gma_state_dicts_load_strictly_both_ways_with_the_reference():
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import make_golden
    make_golden.install_shims()
    import core.FlowFormer.PerCostFormer3.gma as ref_gma
    import stitch_b200
    for heads in (1, 4):
        r_att = ref_gma.Attention(args=None, dim=128, heads=heads, max_pos_size=160, dim_head=128)
        o_att = stitch_b200.gma.Attention(args=None, dim=128, heads=heads, max_pos_size=160, dim_head=128)
        o_att.load_state_dict(r_att.state_dict(), strict=True)
        r_att.load_state_dict(o_att.state_dict(), strict=True)
        r_agg = ref_gma.Aggregate(args=None, dim=128, dim_head=128, heads=heads)
        o_agg = stitch_b200.gma.Aggregate(args=None, dim=128, dim_head=128, heads=heads)
        o_agg.load_state_dict(r_agg.state_dict(), strict=True)
        r_agg.load_state_dict(o_agg.state_dict(), strict=True)


def test_kernels_refuse_autograd_and_cpu_tensors():
    """Inference only, B200 only: a tensor that requires grad (with autograd on) or lives on the CPU raises."""
    import stitch_b200
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        stitch_b200.warp(torch.zeros(1, 3, 8, 8), torch.zeros(1, 2, 8, 8))
    ad = stitch_b200.FlowHomoAdpater(torch.nn.Identity(), torch.nn.Identity(), object())
    with torch.enable_grad():
        with pytest.raises(NotImplementedError, match="inference-only"):
            ad(torch.zeros(1, 3, 8, 8), torch.zeros(1, 3, 8, 8), type="train")
