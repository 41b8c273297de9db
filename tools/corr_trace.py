"""Cost-volume kernel: where each role waits.  Runs a -DSB_CORR_TRACE build
(bash tools/build_variants.sh corr_tcgen05 "trace:-DSB_CORR_TRACE") at batch 16 / 512^2 with the fused pyramid and prints,
per CTA, the cycles the MMA issuer, the first epilogue warp and the producer spent in their waits."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "tools", "probes", "libstitch_corr_tcgen05_trace.so")
os.environ["STITCH_B200_LIB"] = lib
sys.path.insert(0, ROOT)
import torch, stitch_b200 as sb
from stitch_b200 import corr as C
g = torch.Generator(device="cuda").manual_seed(3)
f1, f2 = (torch.randn(16, 256, 64, 64, device="cuda", generator=g) for _ in range(2))
t1, t2 = C.tokens_bf16(f1), C.tokens_bf16(f2)
h = ctypes.CDLL(lib)
if "CORR_TUNE" in os.environ:                       # e.g. CORR_TUNE=14:0 (static unit order)
    k, v = os.environ["CORR_TUNE"].split(":")
    print("sb_tune(%s, %s) ->" % (k, v), sb._lib.load().sb_tune(int(k), int(v)))
for lv in (0, 3):
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); C.corr_from_tokens(t1, t2, 256, (64, 64), (64, 64), pyramid_levels=lv); e1.record()
        torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3
    acc = (ctypes.c_longlong * (148 * 8))()
    assert h.sb_corr_acc_read(acc) == 0
    rows = [[acc[c * 8 + i] for i in range(8)] for c in range(148)]
    loop = sorted(r[5] for r in rows)
    print("pyramid levels=%d: %.1f us; epilogue loop cycles min %d median %d max %d (= %.0f MHz at the slowest)" % (
        lv, us, loop[0], loop[74], loop[-1], loop[-1] / us))
    names = ["MMA: accumulator not drained", "MMA: B stage not loaded", "MMA: A block not loaded",
             "EPI: accumulator not ready", "EPI: staging buffer busy (store path)", "EPI: loop", "PROD: no free B stage"]
    for i, n in enumerate(names):
        v = sorted(r[i] for r in rows)
        print("  %-40s median %8d  min %8d  max %8d   (%.0f %% of the loop)" % (n, v[74], v[0], v[-1], 100.0 * v[74] / loop[74]))
    slow = sorted(range(148), key=lambda c: -rows[c][5])[:5]; fast = sorted(range(148), key=lambda c: rows[c][5])[:5]
    acc2 = (ctypes.c_longlong * (148 * 4))()
    if hasattr(h, "sb_corr_acc2_read") and h.sb_corr_acc2_read(acc2) == 0:
        for i, n in enumerate(["EPI: tcgen05.ld round trips (volume slices)", "EPI: pooling + st.shared", "EPI: fence.proxy.async + TMA store issue"]):
            v = sorted(acc2[c * 4 + i] for c in range(148))
            print("  %-44s median %8d  (%.0f %% of the loop)" % (n, v[74], 100.0 * v[74] / loop[74]))
    for tag, lst in (("slowest", sorted(range(148), key=lambda c: -rows[c][5])[:8]), ("fastest", sorted(range(148), key=lambda c: rows[c][5])[:4])):
        for c in lst:
            r = rows[c]
            print("  %s cta %3d smid %3d loop %7d | MMA wait acc %6d B %6d A %6d | EPI wait acc %6d stage %6d | PROD %6d" % (tag, c, r[7], r[5], r[0], r[1], r[2], r[3], r[4], r[6]))
    print("  slowest CTAs (cta, smid, loop):", [(c, rows[c][7], rows[c][5]) for c in slow], " fastest:", [(c, rows[c][7], rows[c][5]) for c in fast])
