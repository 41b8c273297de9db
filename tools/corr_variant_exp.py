"""Cost-volume kernel variants (tools/build_variants.sh corr_tcgen05 ...): times corr_from_tokens at batch 16 (no pyramid /
3 fused levels) with the default library and with every tools/probes/libstitch_corr_tcgen05_*.so, one subprocess per
library, and checks that the outputs are bit-identical to the default library's."""
import glob, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "--one":
    sys.path.insert(0, ROOT)
    import torch, stitch_b200 as sb
    from stitch_b200 import corr as C
    g = torch.Generator(device="cuda").manual_seed(3)
    f1, f2 = (torch.randn(16, 256, 64, 64, device="cuda", generator=g) for _ in range(2))
    t1, t2 = C.tokens_bf16(f1), C.tokens_bf16(f2)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    line = os.path.basename(os.environ.get("STITCH_B200_LIB", "default"))
    if "CORR_TUNE" in os.environ:                   # e.g. CORR_TUNE=14:0 (static unit order)
        from stitch_b200 import _lib
        k, v = os.environ["CORR_TUNE"].split(":")
        _lib.load().sb_tune(int(k), int(v))
        line += " tune " + os.environ["CORR_TUNE"]
    for lv in (0, 3):
        f = lambda: C.corr_from_tokens(t1, t2, 256, (64, 64), (64, 64), pyramid_levels=lv)
        out = f(); torch.cuda.synchronize()
        ts = []
        for _ in range(12):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); f(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
        ts.sort()
        res = out if lv == 0 else out[0]
        chk = float(res.double().sum().item()), float(res.reshape(16, -1)[3, 100 * 4096:101 * 4096].double().abs().sum().item())
        if lv:
            chk += tuple(float(t.double().sum().item()) for t in out[1])
        line += "  | levels=%d: %.1f us (min), %.1f (median)  checksum %s" % (lv, ts[0] * 1e3, ts[6] * 1e3, "%.6e" % sum(chk))
    print(line, flush=True)
    sys.exit(0)
for lib in [None] + sorted(glob.glob(os.path.join(ROOT, "tools", "probes", "libstitch_corr_tcgen05_*.so"))):
    env = dict(os.environ)
    if lib:
        env["STITCH_B200_LIB"] = lib
    subprocess.run([sys.executable, __file__, "--one"], env=env, timeout=600)
    if lib is None and "CORR_TUNE_ALSO" in os.environ:      # the default library once more with a tune key, e.g. 14:0
        env["CORR_TUNE"] = os.environ["CORR_TUNE_ALSO"]
        subprocess.run([sys.executable, __file__, "--one"], env=env, timeout=600)
