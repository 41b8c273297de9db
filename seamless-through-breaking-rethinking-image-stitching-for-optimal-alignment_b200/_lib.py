"""ctypes binding of libstitchb200.so (the C ABI declared in include/stitch_b200.h).

There is no CPU fallback and no alternative backend: if the shared library is
missing it is built with nvcc; if that is impossible, or a kernel is asked to
run on something that is not a B200, the call raises.
"""
from __future__ import annotations

import ctypes
import os
import threading
from ctypes import c_char_p, c_float, c_int, c_longlong, c_size_t, c_uint, c_void_p

import torch

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
# STITCH_B200_LIB: load another build of the same library (kernel experiments, tools/)
LIB_PATH = os.environ.get("STITCH_B200_LIB") or os.path.join(_PKG_DIR, "lib", "libstitchb200.so")

_lock = threading.Lock()
_lib = None

# name -> (restype, argtypes); mirrors include/stitch_b200.h one to one
_P = c_void_p
SIGNATURES = {
    "sb_version": (c_int, []),
    "sb_last_error": (c_char_p, []),
    "sb_device_check": (c_int, []),
    "sb_launch_count": (c_longlong, []),
    "sb_reset_launch_count": (None, []),
    "sb_shutdown": (c_int, []),
    "sb_debug_word": (c_uint, []),
    "sb_tune": (c_int, [c_int, c_int]),
    "sb_host_alloc": (c_void_p, [c_size_t, c_int]),
    "sb_host_free": (c_int, [c_void_p]),
    "sb_corr_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "sb_corr": (c_int, [_P, _P, _P, _P, _P, _P, _P, c_size_t, c_int, c_int, c_int, c_int, c_int, c_int, _P]),
    "sb_feat_to_tokens_bf16": (c_int, [_P, _P, c_int, c_int, c_int, _P]),
    "sb_corr_tokens": (c_int, [_P, _P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, c_int, _P]),
    "sb_corr_tokens_bidir": (c_int, [_P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, _P]),
    "sb_attn_softmax_tokens": (c_int, [_P, _P, _P, _P, c_int, c_int, c_int, c_int, _P]),
    "sb_corr_tokens_pitched": (c_int, [_P, _P, _P, c_longlong, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, c_int, _P]),
    "sb_corr_tokens_bf16out": (c_int, [_P, _P, _P, c_int, c_int, c_int, c_int, c_int, c_int, _P]),
    "sb_avg_pool2x2": (c_int, [_P, _P, c_longlong, c_int, c_int, _P]),
    "sb_corr_lookup": (c_int, [_P, _P, _P, c_int, c_int, c_int, c_int, c_int, c_int, c_float, c_int, c_int, _P]),
    "sb_bilinear_sampler": (c_int, [_P, _P, _P, c_int, c_int, c_int, c_int, c_int, c_int, _P]),
    "sb_flow_warp": (c_int, [_P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, _P]),
    "sb_flow_warp_nearest": (c_int, [_P, _P, _P, c_int, c_int, c_int, c_int, _P]),
    "sb_homo_warp": (c_int, [_P, _P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, _P]),
    "sb_dlt_theta": (c_int, [_P, _P, _P, _P, _P, _P, _P, c_int, _P]),
    "sb_tps_warp": (c_int, [_P, _P, _P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, _P]),
    "sb_tps_kornia_warp": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, c_int, _P]),
    "sb_grid_sample": (c_int, [_P, _P, _P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, _P]),
    "sb_softmax_rows": (c_int, [_P, c_longlong, c_int, c_longlong, c_int, _P]),
    "sb_attn_aggregate": (c_int, [_P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, _P]),
    "sb_gemm_nt_tf32": (c_int, [_P, _P, _P, c_int, c_int, c_int, c_int, _P]),
    "sb_ccl_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "sb_ccl": (c_int, [_P, _P, _P, _P, c_size_t, c_int, c_int, c_int, c_int, c_float, _P]),
    "sb_softmax_rows_bf16": (c_int, [_P, _P, c_longlong, c_int, c_longlong, _P]),
    "sb_attn_aggregate_bf16": (c_int, [_P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, _P]),
    "sb_upsample_flow": (c_int, [_P, _P, _P, c_int, c_int, c_int, _P]),
    "sb_patch_embed_pack_bytes": (c_size_t, []),
    "sb_patch_embed_pack": (c_int, [_P, _P, _P, _P, _P]),
    "sb_patch_embed_proj": (c_int, [_P, _P, _P, _P, c_longlong, c_int, c_int, _P]),
    "sb_range_map": (c_int, [_P, _P, _P, c_int, c_int, c_int, c_int, _P]),
    "sb_morph_open": (c_int, [_P, _P, c_int, c_int, c_int, c_int, c_int, c_int, _P]),
    "sb_composite_test_out": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, c_int, c_int, c_int, _P]),
    "sb_build_model": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, c_int, c_int, c_int, _P]),
    "sb_tps_mix_blend": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, c_int, c_int, c_int, _P]),
    "sb_overlap_mask": (c_int, [_P, _P, c_int, c_int, c_int, _P]),
}


def load(build_if_missing: bool = True) -> ctypes.CDLL:
    """Load (building first if needed) libstitchb200.so and declare every symbol."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            if not build_if_missing:
                raise RuntimeError(f"{LIB_PATH} is missing (run __graft_entry__.build())")
            from .build import build

            build()
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the .so lacks a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = lib
        return lib


def last_error() -> str:
    msg = load().sb_last_error()
    return msg.decode("utf-8", "replace") if msg else ""


def check(rc: int, what: str) -> None:
    if rc != 0:
        raise RuntimeError(f"{what} failed (code {rc}): {last_error()}")


def stream_ptr() -> c_void_p:
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t: torch.Tensor | None) -> c_void_p:
    if t is None:
        return c_void_p(0)
    return c_void_p(t.data_ptr())


def dev_f32(t: torch.Tensor, name: str) -> torch.Tensor:
    """The kernels take dense fp32 CUDA tensors; anything else is an error for
    the device (no CPU path exists) and a cheap normalisation for layout/dtype."""
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name}: expected a torch.Tensor, got {type(t)}")
    if not t.is_cuda:
        raise RuntimeError(
            f"{name}: tensor is on {t.device}; stitch_b200 runs on B200 GPUs only (no CPU fallback)")
    if t.requires_grad and torch.is_grad_enabled():
        # The kernels are forward-only: they are reached through raw pointers, so autograd would silently see a
        # detached result (a loss without grad_fn, weights that never train).  Inference only — say so loudly.
        raise RuntimeError(
            f"{name}: requires_grad tensor with autograd enabled; stitch_b200 kernels are inference-only "
            "(no backward pass): call under torch.no_grad() / model.eval(), or detach the input")
    if t.device.index != torch.cuda.current_device():
        # kernels launch on the calling thread's current device and stream (as nn.DataParallel's replica threads
        # set them); a tensor that lives elsewhere would be dereferenced on the wrong GPU
        raise RuntimeError(
            f"{name}: tensor is on {t.device} but the current CUDA device is cuda:{torch.cuda.current_device()}; "
            f"wrap the call in `with torch.cuda.device({t.device.index}):`")
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def launch_count() -> int:
    return int(load().sb_launch_count())


def reset_launch_count() -> None:
    load().sb_reset_launch_count()


_HOST_BUFFERS = []   # owners of cudaHostAlloc'd blocks; views of a block may outlive the tensor that
                     # pinned_like() returned, so blocks are only released by free_pinned_buffers()


class _HostBuffer:
    """Owner of one cudaHostAlloc'd block."""

    def __init__(self, nbytes: int, write_combined: bool):
        self.ptr = load().sb_host_alloc(nbytes, 1 if write_combined else 0)
        if not self.ptr:
            raise RuntimeError(f"sb_host_alloc({nbytes}): {last_error()}")
        self.nbytes = nbytes
        self.buf = (ctypes.c_char * nbytes).from_address(self.ptr)

    def free(self):
        if self.ptr:
            load().sb_host_free(self.ptr)
            self.ptr = None


def pinned_like(t: torch.Tensor, write_combined: bool = False) -> torch.Tensor:
    """A page-locked host tensor with ``t``'s shape / dtype holding a copy of ``t`` (CPU tensor)."""
    src = t.contiguous()
    nbytes = max(src.numel() * src.element_size(), 1)
    owner = _HostBuffer(nbytes, write_combined)
    _HOST_BUFFERS.append(owner)
    out = torch.frombuffer(owner.buf, dtype=src.dtype, count=src.numel()).view(src.shape)
    out.copy_(src)
    return out


def free_pinned_buffers() -> None:
    """Release every block handed out by pinned_like(); the caller guarantees no tensor uses them any more."""
    while _HOST_BUFFERS:
        _HOST_BUFFERS.pop().free()
