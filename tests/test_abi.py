"""CPU: the C-ABI shared library loads and exports every symbol that
include/stitch_b200.h declares; compute entry points refuse to run without a B200."""
import ctypes
import os
import re

import pytest

from conftest import ROOT


@pytest.fixture(scope="module")
def lib():
    import stitch_b200
    return stitch_b200._lib.load()


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "stitch_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(sb_[a-z0-9_]+)\s*\(", hdr)))


def test_header_symbols_exported(lib):
    import stitch_b200
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/stitch_b200.h but not exported"
    # and the ctypes table binds exactly the declared set
    assert sorted(stitch_b200._lib.SIGNATURES) == names


def test_version_and_error_string(lib):
    assert lib.sb_version() == 100
    assert isinstance(lib.sb_last_error(), (bytes, type(None)))


def test_sm100a_only_binary():
    import subprocess
    import stitch_b200
    try:
        out = subprocess.run(["cuobjdump", "-lelf", stitch_b200._lib.LIB_PATH], capture_output=True, text=True).stdout
    except FileNotFoundError:
        pytest.skip("cuobjdump not available")
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_no_cpu_fallback(lib):
    import torch
    import stitch_b200
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu tests")
    # the ABI itself reports the missing device instead of computing anything
    assert lib.sb_device_check() != 0
    assert lib.sb_flow_warp(None, None, None, None, None, 1, 1, 1, 1, None) != 0
    assert b"" != lib.sb_last_error()
    # and the Python surface refuses CPU tensors outright
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        stitch_b200.warp(torch.zeros(1, 6, 8, 8), torch.zeros(1, 2, 8, 8))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        stitch_b200.corr_volume(torch.zeros(1, 64, 4, 4), torch.zeros(1, 64, 4, 4))


def test_product_does_not_import_oracle():
    """The product path must never route through the oracle."""
    pkg = os.path.join(ROOT, "seamless-through-breaking-rethinking-image-stitching-for-optimal-alignment_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "stitch_oracle" not in src and "liboracle" not in src, f


def test_patch_reference_assigns_call_sites():
    """Where the reference is mounted (the build container), patch_reference() must replace
    exactly the call-site attributes of SURVEY §8(b). The GPU box has no reference: skipped."""
    import sys
    ref = os.environ.get("STITCH_REFERENCE", "/root/reference")
    if not os.path.isdir(os.path.join(ref, "core")):
        pytest.skip("reference not mounted")
    sys.path[:0] = [os.path.join(ROOT, "tests", "golden")]
    import make_golden
    saved = dict(sys.modules)
    saved_path = list(sys.path)
    try:
        make_golden.install_shims()
        import stitch_b200
        done = stitch_b200.patch_reference()
        import core.warp_utils as wu
        import core.udis_utils.torch_homo_transform as th
        from core.FlowFormer.PerCostFormer3.encoder import MemoryEncoder
        from core.FlowFormer.PerCostFormer3.decoder import MemoryDecoder
        assert wu.warp is stitch_b200.warp_utils.warp
        assert th.transformer is stitch_b200.torch_homo_transform.transformer
        assert MemoryEncoder.corr is stitch_b200.corr.memory_encoder_corr
        assert MemoryDecoder.encode_flow_token is stitch_b200.lookup.memory_decoder_encode_flow_token
        assert MemoryDecoder.upsample_flow is stitch_b200.decoder.memory_decoder_upsample_flow
        import core.FlowFormer.PerCostFormer3.gma as ref_gma
        assert ref_gma.Attention.forward is stitch_b200.gma.attention_forward
        assert ref_gma.Aggregate.forward is stitch_b200.gma.aggregate_forward
        from core.UDIS2.Homography.network import UDIS2Network
        assert UDIS2Network.CCL is stitch_b200.udis2_homography.udis2_network_ccl
        from core.FlowFormer.PerCostFormer3.encoder import PatchEmbed
        assert PatchEmbed.forward is stitch_b200.encoder.patch_embed_forward              # N4
        assert not any(d.endswith("get_tps_transform") for d in done)                   # stays the reference's binding
        assert len(done) >= 17
    finally:
        for k in list(sys.modules):
            if k not in saved and (k.startswith("core") or k.startswith("timm") or k.startswith("skimage")):
                del sys.modules[k]
        sys.path[:] = saved_path


def test_header_is_plain_c():
    """include/stitch_b200.h is the drop-in boundary: it must compile as C99 and as C++17 on its own."""
    import subprocess
    import tempfile
    src = '#include "stitch_b200.h"\nint main(void) { return sb_version() > 0 ? 0 : 1; }\n'
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "hdr.c")
        open(path, "w").write(src)
        inc = os.path.join(ROOT, "include")
        subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-I", inc, path], check=True)
        subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-x", "c++", "-I", inc, path], check=True)


def test_package_level_corr_is_callable_and_a_module():
    """`corr(fmap1, fmap2, heads=1)` at package level; `stitch_b200.corr` is also the module."""
    import inspect
    import torch
    import stitch_b200 as sb
    assert callable(sb.corr) and inspect.ismodule(sb.corr)
    assert list(inspect.signature(sb.corr.corr).parameters)[:3] == ["fmap1", "fmap2", "heads"]
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        sb.corr(torch.zeros(1, 8, 4, 4), torch.zeros(1, 8, 4, 4))


def test_tune_knob_enum_is_consistent():
    """SB_TUNE_* keys are unique, below SB_TUNE_COUNT, and sb_tune rejects anything else."""
    hdr = open(os.path.join(ROOT, "include", "stitch_b200.h")).read()
    keys = dict((m.group(1), int(m.group(2))) for m in re.finditer(r"\b(SB_TUNE_[A-Z0-9_]+)\s*=\s*(\d+)", hdr))
    count = keys.pop("SB_TUNE_COUNT")
    assert len(keys) >= 10 and len(set(keys.values())) == len(keys) and max(keys.values()) < count
    import stitch_b200 as sb
    lib = sb._lib.load()
    for v in keys.values():
        assert lib.sb_tune(v, 0) == 0
    assert lib.sb_tune(count, 1) != 0 and lib.sb_tune(-1, 1) != 0
    assert b"unknown key" in lib.sb_last_error()
