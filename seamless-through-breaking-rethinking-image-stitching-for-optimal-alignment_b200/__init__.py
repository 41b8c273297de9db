"""stitch_b200 — B200-native (sm_100a) kernels for the stitching-alignment hot
path of gargatik/Seamless-Through-Breaking-Rethinking-Image-Stitching-for-Optimal-Alignment,
behind the reference's own Python call surface.

Importing the package needs only torch; the shared library (and a B200) is
needed when a kernel is called. There is no CPU fallback.
"""
from . import _lib  # noqa: F401
from . import (adapter, composition, corr, decoder, encoder, gma, kornia_tps, lookup, patch, pipeline, torch_DLT,  # noqa: F401
               torch_homo_transform, torch_tps_transform, udis2_homography, warp_utils)
from .adapter import FlowHomoAdpater  # noqa: F401
from .composition import build_model, composite_test_out, preprocess_occlusion_mask  # noqa: F401
from .corr import corr as corr_volume, corr_pyramid  # noqa: F401
from .lookup import bilinear_sampler, encode_flow_token, encode_flow_token_pyramid  # noqa: F401
from .patch import patch_reference  # noqa: F401
from .warp_utils import compute_occlusion, compute_range_map, warp  # noqa: F401

__version__ = "0.1.0"
