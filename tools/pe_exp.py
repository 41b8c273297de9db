"""PatchEmbed kernel variants (tools/pe_variants.sh): time sb_patch_embed_proj on 65 536 maps with each library in
tools/probes/libstitch_pe_*.so (one subprocess per library: STITCH_B200_LIB is read at import).  `--check` also compares
with the default library's output (variants that keep the arithmetic must be bit-identical)."""
import glob, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "--one":
    sys.path.insert(0, ROOT)
    import torch, stitch_b200 as sb
    g = torch.Generator(device="cuda").manual_seed(7)
    rnd = lambda *s: torch.randn(*s, device="cuda", generator=g)
    nq = 65536
    maps = rnd(nq, 1, 64, 64) * 8
    w1, b1, w2, b2, w3, b3 = rnd(16, 1, 6, 6) / 6, rnd(16) / 4, rnd(32, 16, 6, 6) / 24, rnd(32) / 4, rnd(64, 32, 6, 6) / 34, rnd(64) / 4
    pack = sb.encoder.pack_patch_embed_weights(w1, w2, w3)
    f = lambda: sb.encoder.patch_embed_proj(maps, w1, b1, w2, b2, w3, b3, pack=pack)
    out = f(); torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ts.sort()
    clk = ""
    try:                                            # SM clock / power while the kernel runs back to back (~0.7 s)
        import pynvml, threading, time
        pynvml.nvmlInit(); hd = pynvml.nvmlDeviceGetHandleByIndex(0)
        samples, stop = [], threading.Event()
        def poll():
            while not stop.is_set():
                samples.append((pynvml.nvmlDeviceGetClockInfo(hd, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetPowerUsage(hd) / 1e3,
                                pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(hd)))
                time.sleep(0.02)
        th = threading.Thread(target=poll); th.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(300): f()
        e1.record(); torch.cuda.synchronize(); stop.set(); th.join()
        sm = sorted(x[0] for x in samples[5:]); pw = sorted(x[1] for x in samples[5:])
        reasons = 0
        for x in samples[5:]: reasons |= x[2]
        clk = "  | 300 back to back: %.1f us each, SM %d MHz median (min %d), %.0f W median (max %.0f), throttle mask 0x%x" % (
            e0.elapsed_time(e1) / 300 * 1e3, sm[len(sm) // 2], sm[0], pw[len(pw) // 2], pw[-1], reasons)
    except Exception as ex:
        clk = "  | clocks unavailable: %r" % (ex,)
    chk = ""
    if len(sys.argv) > 2:
        ref = torch.load(sys.argv[2]) if os.path.exists(sys.argv[2]) else None
        if ref is None:
            torch.save(out[:4096].cpu(), sys.argv[2])
        else:
            chk = " identical" if torch.equal(ref, out[:4096].cpu()) else " DIFFERENT (max %.3g)" % (ref - out[:4096].cpu()).abs().max().item()
    print("%-44s %8.1f us (min of 5, median %.1f)%s" % (os.path.basename(os.environ.get("STITCH_B200_LIB", "default")), ts[0] * 1e3, ts[2] * 1e3, chk) + clk, flush=True)
    sys.exit(0)
libs = [None] + sorted(glob.glob(os.path.join(ROOT, "tools", "probes", "libstitch_pe_*.so")) + glob.glob(os.path.join(ROOT, "tools", "probes", "libstitch_patch_embed_*.so")))
for lib in libs:
    env = dict(os.environ)
    if lib:
        env["STITCH_B200_LIB"] = lib
    subprocess.run([sys.executable, __file__, "--one", "/tmp/pe_ref.pt"], env=env, timeout=600)
