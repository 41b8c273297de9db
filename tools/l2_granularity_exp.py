"""Experiment: effect of cudaLimitMaxL2FetchGranularity (32/64/128 B) on the gather kernels."""
import ctypes, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import stitch_b200 as sb
from stitch_b200 import corr as C
from kernel_bench import timeit

rt = ctypes.CDLL("libcudart.so.12")
LIMIT = 5  # cudaLimitMaxL2FetchGranularity
B, S = 16, 512
g = torch.Generator(device="cuda").manual_seed(0)
rnd = lambda *s: torch.randn(*s, device="cuda", generator=g)
f1, f2 = rnd(B, 256, 64, 64), rnd(B, 256, 64, 64)
t1, t2 = C.tokens_bf16(f1), C.tokens_bf16(f2)
vol = C.corr_from_tokens(t1, t2, 256, (64, 64), (64, 64))
maps = vol.view(B * 4096, 1, 64, 64)
coords = sb.lookup.coords_grid(B, 64, 64, device="cuda") + rnd(B, 2, 64, 64) * 2
x6 = torch.rand(B, 6, S, S, device="cuda", generator=g) * 255
img = torch.rand(B, 3, S, S, device="cuda", generator=g) * 255
flo = torch.nn.functional.interpolate(rnd(B, 2, 64, 64) * 2, size=(S, S), mode="bilinear", align_corners=True)
src = sb.torch_DLT.corner_points(S, S, B, "cuda")
M = sb.torch_DLT.norm_matrix(S / 8, S / 8)
H, th, thi = sb.torch_DLT.dlt_thetas(src / 8, (src + rnd(B, 4, 2) * 20) / 8, left=sb.torch_DLT._inv3(M), right=M)
for gran in (64, 32, 128, 64):
    v = ctypes.c_size_t(0)
    rc = rt.cudaDeviceSetLimit(LIMIT, ctypes.c_size_t(gran))
    rt.cudaDeviceGetLimit(ctypes.byref(v), LIMIT)
    torch.cuda.synchronize()
    res = {
        "lookup": timeit(lambda: sb.encode_flow_token(maps, coords)),
        "flow_warp": timeit(lambda: sb.warp(x6, flo)),
        "homo3+1": timeit(lambda: sb.torch_homo_transform.transformer(img, th, (S, S), append_ones=3)),
        "range": timeit(lambda: sb.compute_occlusion(flo, flo, "wang", occlusion_are_zeros=True, threshold=True)),
        "corr_lv3": timeit(lambda: C.corr_from_tokens(t1, t2, 256, (64, 64), (64, 64), pyramid_levels=3)),
        "tokens": timeit(lambda: C.tokens_bf16(f1)),
    }
    print(f"granularity set rc={rc} -> {v.value} B: " + "  ".join(f"{k} {ms*1e3:.1f}us" for k, ms in res.items()), flush=True)


# ---- which pooled level costs what in the fused epilogue
lib = sb._lib.load()
P = sb._lib.ptr
n = 4096
volb = torch.empty((B, n, n), device="cuda")
l1 = torch.empty((B * n, 32, 32), device="cuda")
l2 = torch.empty((B * n, 16, 16), device="cuda")
l3 = torch.empty((B * n, 8, 8), device="cuda")
def run(a, b_, c):
    sb._lib.check(lib.sb_corr_tokens(P(t1), P(t2), P(volb), P(a), P(b_), P(c), B, 256, 64, 64, 64, 64, sb._lib.stream_ptr()), "corr")
for name, args in (("none", (None, None, None)), ("l1", (l1, None, None)), ("l1+l2", (l1, l2, None)), ("l1+l2+l3", (l1, l2, l3)),
                   ("l2+l3", (None, l2, l3))):
    print(f"corr levels {name:9s}: {timeit(lambda: run(*args))*1e3:.1f} us", flush=True)
