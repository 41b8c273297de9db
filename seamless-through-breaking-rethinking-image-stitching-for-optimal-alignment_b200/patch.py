"""patch_reference() — drop the B200 kernels in behind the reference's own call
surface by assigning this package's functions over the reference modules'
attributes (SURVEY §8b lists the call sites). The reference must already be
importable (``sys.path`` containing its root and ``core/``).

INFERENCE ONLY.  The kernels have no backward pass: run the patched reference the way its own
``evaluate.py`` / ``out.py`` do — ``model.eval()`` under ``torch.no_grad()``.  With autograd enabled any
patched function that receives a tensor requiring grad raises (it does not return a detached result)."""
from __future__ import annotations

import importlib

__all__ = ["patch_reference"]


def patch_reference(verbose: bool = False):
    from . import composition, corr, decoder, encoder, gma, kornia_tps, lookup, udis2_homography, torch_homo_transform, torch_tps_transform, warp_utils

    done = []

    def _set(modname, attr, fn, cls=None):
        try:
            mod = importlib.import_module(modname)
        except Exception as e:  # the reference module itself may be un-importable (missing timm, ...)
            if verbose:
                print(f"[stitch_b200] skip {modname}.{attr}: {e}")
            return
        target = getattr(mod, cls) if cls else mod
        setattr(target, attr, fn)
        done.append(f"{modname}.{(cls + '.') if cls else ''}{attr}")

    _set("core.warp_utils", "warp", warp_utils.warp)
    _set("core.warp_utils", "compute_range_map", warp_utils.compute_range_map)
    _set("core.warp_utils", "compute_occlusion", warp_utils.compute_occlusion)
    _set("core.udis_utils.torch_homo_transform", "transformer", torch_homo_transform.transformer)
    _set("core.udis_utils.torch_tps_transform", "transformer", torch_tps_transform.transformer)
    _set("core.udis_utils.torch_tps_transform2", "transformer", torch_tps_transform.transformer)
    _set("core.utils.utils", "bilinear_sampler", lookup.bilinear_sampler)
    _set("core.UDIS2.Composition.network", "build_model", composition.build_model)
    _set("core.flowHomoAdpater", "preprocess_occlusion_mask", composition.preprocess_occlusion_mask)
    _set("core.flowHomoAdpater", "warp", warp_utils.warp)                 # star-imported name (:16)
    _set("core.flowHomoAdpater", "compute_occlusion", warp_utils.compute_occlusion)
    _set("core.FlowFormer.PerCostFormer3.encoder", "corr", corr.memory_encoder_corr, cls="MemoryEncoder")
    _set("core.FlowFormer.PerCostFormer3.encoder", "forward", encoder.patch_embed_forward, cls="PatchEmbed")
    _set("core.FlowFormer.PerCostFormer3.decoder", "encode_flow_token",
         lookup.memory_decoder_encode_flow_token, cls="MemoryDecoder")
    _set("core.FlowFormer.PerCostFormer3.decoder", "upsample_flow",
         decoder.memory_decoder_upsample_flow, cls="MemoryDecoder")
    _set("core.FlowFormer.PerCostFormer3.gma", "forward", gma.attention_forward, cls="Attention")
    _set("core.FlowFormer.PerCostFormer3.gma", "forward", gma.aggregate_forward, cls="Aggregate")
    _set("core.UDIS2.Homography.network", "CCL", udis2_homography.udis2_network_ccl, cls="UDIS2Network")
    # tps_method="kornia" branch (tps_pipline.py:364-381 imports the name at call time).  Only the dense warp is
    # a kernel; `get_tps_transform` stays what the reference binds it to (kornia's own solve, kornia_tps.py:1):
    # replacing it with the pinverse variant the reference defines but never calls would change the numerics
    # of an often ill-conditioned system.
    _set("core.inference.tps_methods.kornia_tps", "warp_image_tps", kornia_tps.warp_image_tps)
    return done
