"""Per-kernel device timings at BASELINE config-2 size (B=16, 512^2), CUDA events,
inputs larger than L2 or rotated so nothing is re-served from L2. Prints achieved
algorithmic GB/s against MEASURED_PEAKS.json."""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import stitch_b200 as sb
from stitch_b200 import corr as C

PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0


def timeit(fn, n=20, warm=3):
    for _ in range(warm):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def report(name, ms, nbytes):
    gbs = nbytes / ms / 1e6
    print(f"{name:34s} {ms*1e3:9.1f} us  {nbytes/1e6:9.1f} MB  {gbs:8.0f} GB/s  {100*gbs/PEAK:5.1f}% of HBM peak", flush=True)


def main():
    B, S = 16, 512
    g = torch.Generator(device="cuda").manual_seed(0)
    rnd = lambda *s: torch.randn(*s, device="cuda", generator=g)
    # ---- corr
    f1, f2 = rnd(B, 256, 64, 64), rnd(B, 256, 64, 64)
    report("feat_to_tokens_bf16", timeit(lambda: C.tokens_bf16(f1)), B * 4096 * 256 * 6)
    t1, t2 = C.tokens_bf16(f1), C.tokens_bf16(f2)
    n = 4096
    for lv in (0, 3):
        ms = timeit(lambda: C.corr_from_tokens(t1, t2, 256, (64, 64), (64, 64), pyramid_levels=lv))
        byt = B * (n * n * 4 * (1 + (0.328125 if lv else 0)) + 2 * n * 256 * 2)
        report(f"corr_umma (pyramid levels={lv})", ms, byt)
        print(f"{'':34s} tensor: {B*2*n*n*256/ms/1e9:.0f} TFLOP/s")
    ms = timeit(lambda: C.corr_from_tokens(t1, t2, 256, (64, 64), (64, 64), out_dtype=torch.bfloat16))
    report("corr_umma (bf16 volume, opt-in)", ms, B * (n * n * 2 + 2 * n * 256 * 2))
    TP = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    tf = B * 2 * n * n * 256 / ms / 1e9
    print(f"{'':34s} tensor: {tf:.0f} TFLOP/s = {100*tf/TP.get('bf16_tflops_sustained', 1404.5):.0f}% of sustained / {100*tf/TP.get('bf16_tflops', 1672.2):.0f}% of burst bf16 peak")
    # ---- lookup
    vol = C.corr_from_tokens(t1, t2, 256, (64, 64), (64, 64))
    maps = vol.view(B * n, 1, 64, 64)
    coords = sb.lookup.coords_grid(B, 64, 64, device="cuda") + rnd(B, 2, 64, 64) * 2
    report("corr_lookup r=4", timeit(lambda: sb.encode_flow_token(maps, coords)), B * n * 732)
    pyr = sb.corr_pyramid(f1, f2, 4)
    report("corr_lookup pyramid (4 levels)", timeit(lambda: sb.encode_flow_token_pyramid(pyr, coords)), B * n * 2904)
    del vol, maps, pyr
    # ---- warps
    img = torch.rand(B, 3, S, S, device="cuda", generator=g) * 255
    x6 = torch.rand(B, 6, S, S, device="cuda", generator=g) * 255
    lo = rnd(B, 2, 64, 64) * 2
    flo = torch.nn.functional.interpolate(lo, size=(S, S), mode="bilinear", align_corners=True)
    occ = (torch.rand(B, 1, S, S, device="cuda", generator=g) < 0.8).float()
    px = B * S * S
    report("flow_warp C=6", timeit(lambda: sb.warp(x6, flo)), px * 56)
    report("flow_warp C=6 +mask +overlap", timeit(lambda: sb.warp(x6, flo, mul_mask=occ, return_overlap=True)), px * 64)
    noise = rnd(B, 2, S, S) * 4
    report("flow_warp C=6 (noise flow)", timeit(lambda: sb.warp(x6, noise)), px * 56)
    src = sb.torch_DLT.corner_points(S, S, B, "cuda")
    M = sb.torch_DLT.norm_matrix(S / 8, S / 8)
    H, th, thi = sb.torch_DLT.dlt_thetas(src / 8, (src + rnd(B, 4, 2) * 20) / 8, left=sb.torch_DLT._inv3(M), right=M)
    report("dlt_theta", timeit(lambda: sb.torch_DLT.dlt_thetas(src / 8, src / 8 + 1, left=sb.torch_DLT._inv3(M), right=M)), B * 200)
    report("homo_warp C=6", timeit(lambda: sb.torch_homo_transform.transformer(x6, th, (S, S))), px * 48)
    report("homo_warp C=3 + 3 ones", timeit(lambda: sb.torch_homo_transform.transformer(img, th, (S, S), append_ones=3)), px * 36)
    report("range_map (occlusion, thresholded)", timeit(lambda: sb.compute_occlusion(flo, flo, "wang", occlusion_are_zeros=True, threshold=True)), px * 28)
    report("morph_open 19x19", timeit(lambda: sb.preprocess_occlusion_mask(occ)), px * 8)
    h1, h2, fw = x6, x6.flip(0), x6.flip(1)
    report("composite_test_out", timeit(lambda: sb.composite_test_out(h1, h2, fw, occ)), px * 119)
    net_out = torch.rand(B, 1, S, S, device="cuda", generator=g)
    report("build_model", timeit(lambda: sb.build_model(lambda *a: net_out, img, img, img, img)), px * 88)
    # ---- GMA (next row 1): attention once, aggregate per GRU iteration
    fm = rnd(B, 128, 64, 64)
    w_qk, w_v, gam = rnd(256, 128, 1, 1) * 0.02, rnd(128, 128, 1, 1) * 0.09, torch.tensor([0.5], device="cuda")
    attn = sb.gma.attention(fm, w_qk)
    sim = torch.empty_like(attn)
    def _softmax_of_copy():
        sim.copy_(attn)
        sb.gma.softmax_rows_(sim)
    report("gma softmax_rows (in place)", timeit(_softmax_of_copy, n=5) - timeit(lambda: sim.copy_(attn), n=5), B * n * n * 8)
    qq, kk = sb.gma.project_qk(fm, w_qk)
    tq, tk = sb.corr.tokens_bf16(qq), sb.corr.tokens_bf16(kk)
    stats = torch.empty(B, n, 2, device="cuda")
    lib = sb._lib.load()
    P = sb._lib.ptr
    def _fused():
        sb._lib.check(lib.sb_attn_softmax_tokens(P(tq), P(tk), P(sim), P(stats), B, 128, n, n, sb._lib.stream_ptr()), "attn")
    ms = timeit(_fused, n=5)
    report("gma attention, fused 2-pass softmax", ms, B * n * n * 4)
    print(f"{'':34s} (unfused: q.k^T volume + softmax_rows = the two lines above; 2 x {B*2*n*n*128/1e9:.0f} GFLOP recomputed)")
    vv = torch.nn.functional.conv2d(fm, w_v).view(B, 128, n)
    ms = timeit(lambda: sb.gma.attn_matmul_v(attn.view(B, n, n), vv, residual=fm.view(B, 128, n), gamma=gam), n=10)
    report("gma attn @ v (tf32 tcgen05)", ms, B * (n * n * 4 + 3 * n * 128 * 4))
    print(f"{'':34s} tensor: {B*2*n*n*128/ms/1e9:.0f} TFLOP/s (tf32)")
    attn16 = attn.view(B, n, n).bfloat16()
    ms = timeit(lambda: sb.gma.attn_matmul_v(attn16, vv, residual=fm.view(B, 128, n), gamma=gam), n=10)
    report("gma attn @ v (bf16 attn opt-in)", ms, B * (n * n * 2 + n * 128 * 2 + 2 * n * 128 * 4))
    del attn16
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    report("  torch fp32 bmm (reference path)", timeit(lambda: torch.bmm(attn.view(B, n, n), vv.transpose(1, 2)), n=3), B * (n * n * 4 + 2 * n * 128 * 4))
    torch.backends.cuda.matmul.allow_tf32 = prev
    del attn, sim
    # ---- CCL (next row 3): [B, 1024, 32, 32] features
    cf1, cf2 = torch.relu(rnd(B, 1024, 32, 32)), torch.relu(rnd(B, 1024, 32, 32))
    ms = timeit(lambda: sb.udis2_homography.CCL(cf1, cf2), n=10)
    report("CCL (norm + tf32 corr + softmax-flow)", ms, B * (2 * 1024 * 1024 * 4 * 3 + 1024 * 1024 * 4 * 2))
    print(f"{'':34s} {B*2*1024*1024*1024/ms/1e9:.0f} TFLOP/s on the contraction actually computed; reference form = 9x the FLOPs")
    # ---- PatchEmbed projection (next row 4): one direction's 65536 cost maps
    if os.environ.get("SB_BENCH_PATCH_EMBED", "1") == "1":
        vol4 = C.corr_from_tokens(t1, t2, 256, (64, 64), (64, 64)).view(B * n, 1, 64, 64)
        gw = lambda o, c: ((torch.rand(o, c, 6, 6, device="cuda", generator=g) * 2 - 1) / (c * 36) ** 0.5, (torch.rand(o, device="cuda", generator=g) * 2 - 1) / (c * 36) ** 0.5)
        (pw1, pb1), (pw2, pb2), (pw3, pb3) = gw(16, 1), gw(32, 16), gw(64, 32)
        pack = sb.encoder.pack_patch_embed_weights(pw1, pw2, pw3)
        ms = timeit(lambda: sb.encoder.patch_embed_proj(vol4, pw1, pb1, pw2, pb2, pw3, pb3, pack=pack), n=5, warm=2)
        report("patch_embed proj (3 convs, tcgen05)", ms, B * n * 2 * 16384)
        fl = B * n * 2.0 * (1024 * 16 * 36 + 256 * 32 * 576 + 64 * 64 * 1152)
        print(f"{'':34s} tensor: {fl/ms/1e9:.0f} TFLOP/s useful = {100*fl/ms/1e9/TP.get('bf16_tflops_sustained', 1404.5):.0f}% of sustained bf16 peak")
        x8 = vol4[:8192]
        def _torch_ref():
            a = torch.relu(torch.nn.functional.conv2d(x8, pw1, pb1, stride=2, padding=2))
            a = torch.relu(torch.nn.functional.conv2d(a, pw2, pb2, stride=2, padding=2))
            return torch.nn.functional.conv2d(a, pw3, pb3, stride=2, padding=2)
        ms8 = timeit(_torch_ref, n=3, warm=1)
        print(f"{'':34s} torch/cuDNN fp32 convs (the reference's path) on 8192 maps: {ms8*1e3:.0f} us -> {ms8*8*1e3:.0f} us per 65536 maps")
        del vol4, x8
    lo2, um = rnd(B, 2, 64, 64), rnd(B, 576, 64, 64)
    report("upsample_flow (convex 8x)", timeit(lambda: sb.decoder.upsample_flow(lo2, um)), B * 4096 * (576 + 128 + 2) * 4)
    ys, xs = torch.meshgrid(torch.linspace(-1, 1, 13, device="cuda"), torch.linspace(-1, 1, 13, device="cuda"), indexing="ij")
    sp = torch.stack([xs, ys], -1).reshape(1, -1, 2).repeat(B, 1, 1)
    tg = sp + 0.02 * rnd(B, 169, 2)
    T = sb.torch_tps_transform.solve_system(sp, tg)
    lib = sb._lib.load()
    xs_t, ys_t = sb.torch_homo_transform.linspace_table(S, x6.device), sb.torch_homo_transform.linspace_table(S, x6.device)
    out = torch.empty_like(x6)
    def tps_only():
        sb._lib.check(lib.sb_tps_warp(sb._lib.ptr(x6), sb._lib.ptr(T), sb._lib.ptr(sp), sb._lib.ptr(xs_t), sb._lib.ptr(ys_t),
                                      sb._lib.ptr(out), None, B, 6, S, S, S, S, 169, sb._lib.stream_ptr()), "tps")
    ms = timeit(tps_only, n=5)
    report("tps_warp pn=169 (kernel only)", ms, px * 48)
    print(f"{'':34s} {px*169/ms/1e6:.1f} G basis evaluations/s")


if __name__ == "__main__":
    main()
